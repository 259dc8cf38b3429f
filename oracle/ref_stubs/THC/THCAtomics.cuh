// Stub: on sm_100a the CUDA toolkit's own atomicAdd overloads (float, double) are what
// THCAtomics.cuh forwards to for the types the reference kernels instantiate.
#pragma once
#include <cuda_runtime.h>
