// nccl.h (CPU EMULATION SHIM) - TEST INFRASTRUCTURE ONLY: one rank, every collective is an error.
#pragma once
#include <string.h>
typedef struct emuNcclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclInvalidUsage = 5 };
enum ncclDataType_t { ncclDouble = 8 };
enum ncclRedOp_t { ncclSum = 0 };
static inline const char *ncclGetErrorString(ncclResult_t) { return "NCCL is not available in the CPU emulation (single rank only)"; }
static inline ncclResult_t ncclGetUniqueId(ncclUniqueId *u) { memset(u, 0, sizeof *u); return ncclSuccess; }
static inline ncclResult_t ncclCommInitRank(ncclComm_t *, int, ncclUniqueId, int) { return ncclInvalidUsage; }
static inline ncclResult_t ncclCommDestroy(ncclComm_t) { return ncclSuccess; }
static inline ncclResult_t ncclAllReduce(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
static inline ncclResult_t ncclSend(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
static inline ncclResult_t ncclRecv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
static inline ncclResult_t ncclGroupStart() { return ncclSuccess; }
static inline ncclResult_t ncclGroupEnd() { return ncclSuccess; }
