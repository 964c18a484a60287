"""Reader for the record files written by oracle/ref_dump.c and oracle/martini_oracle (test infrastructure).

Record = name[32] | dtype char ('d' f64, 'q' u64, 'i' i32) | pad[7] | count u64 | payload.
"""
import numpy as np

_DT = {"d": np.float64, "q": np.uint64, "i": np.int32}


def read_records(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        name = data[off:off + 32].split(b"\0")[0].decode()
        dt = _DT[chr(data[off + 32])]
        count = int(np.frombuffer(data, np.uint64, 1, off + 40)[0])
        off += 48
        nbytes = count * np.dtype(dt).itemsize
        out[name] = np.frombuffer(data, dt, count, off).copy()
        off += nbytes
    return out


def write_records(path, recs):
    inv = {np.dtype(np.float64): b"d", np.dtype(np.uint64): b"q", np.dtype(np.int32): b"i"}
    with open(path, "wb") as f:
        for name, arr in recs.items():
            arr = np.ascontiguousarray(arr)
            hdr = name.encode()[:31].ljust(32, b"\0") + inv[arr.dtype] + b"\0" * 7
            f.write(hdr)
            f.write(np.uint64(arr.size).tobytes())
            f.write(arr.tobytes())
