// stand-in for the driver API header: only the types wspr_decode.cu names for its (optional) green-context partition
#pragma once
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
typedef int CUdevice;
typedef struct CUgreenCtx_st *CUgreenCtx;
typedef struct CUstream_st *CUstream;
typedef struct CUdevResourceDesc_st *CUdevResourceDesc;
typedef enum { CU_DEV_RESOURCE_TYPE_INVALID = 0, CU_DEV_RESOURCE_TYPE_SM = 1 } CUdevResourceType;
typedef struct {
    CUdevResourceType type;
    struct { unsigned smCount; } sm;
} CUdevResource;
enum { CU_GREEN_CTX_DEFAULT_STREAM = 1, CU_STREAM_NON_BLOCKING = 1, CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING = 1 };
