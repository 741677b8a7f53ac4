#pragma once
#include "cuda_runtime.h"
#define CUDART_NAN_F __int_as_float(0x7fffffff)
#define CUDART_INF_F __int_as_float(0x7f800000)
