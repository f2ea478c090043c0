/* tests/emu/evp/nccl.h -- the NCCL TYPES evp_halo.cu names (the functions are resolved with dlopen at run time and are
 * never reached by the single-rank emulation).  TEST INFRASTRUCTURE ONLY. */
#ifndef TESTS_EMU_EVP_NCCL_H
#define TESTS_EMU_EVP_NCCL_H
#include <stddef.h>
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3 } ncclResult_t;
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat = 7, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3, ncclAvg = 4 } ncclRedOp_t;
#endif
