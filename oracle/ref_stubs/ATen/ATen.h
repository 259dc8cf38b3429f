// Stub: the reference kernel header includes <ATen/ATen.h> but uses nothing from it.
// (oracle/build_ref.py compiles the REFERENCE's ms_deform_im2col_cuda.cuh without torch headers.)
#pragma once
#include <cstdint>
