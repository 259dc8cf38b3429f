// Stub, see ../ATen.h
#pragma once
#include <cuda_runtime.h>
