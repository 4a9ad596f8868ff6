// TEST INFRASTRUCTURE (oracle) -- not product code.
//
// Prelude force-included in front of the UNMODIFIED reference translation units
// (/root/reference/submodules/diff-gaussian-rasterization/cuda_rasterizer/*.cu) when they
// are compiled, from where they lie, into oracle/_ref/libgsref.so (see build_ref.sh).
//
// The reference ships with heavy debug instrumentation (SURVEY.md header): device printf
// per Gaussian (forward.cu:209,221-356,389; backward.cu:221,347-419,528,584-589), host
// std::cout of whole matrices and per-call D2H copies + file dumps ./means3D_debug.txt,
// ./means2D_backward_debug.txt (rasterizer_impl.cu:227-267,451-463).  None of it changes
// arithmetic.  No reference source is copied or edited; instead every header the reference
// includes is included here FIRST (so its include guard is spent), and then the debug sinks
// are redirected by macros:
//   printf(...)        -> nothing
//   std::cout          -> std::gsref_null_out   (swallows operator<<)
//   std::ofstream      -> std::gsref_null_ofs   (no file is created)
//   cudaMemcpy(...)    -> gsref_memcpy: only the 4-byte num_rendered read
//                         (rasterizer_impl.cu:331) is executed, the debug copies are skipped.
// Also supplies <cstdint>, which rasterizer_impl.h:24 needs under gcc 13.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdio.h>
#include <math.h>
#include <cmath>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <numeric>
#include <vector>
#include <functional>
#include <stdexcept>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_runtime_api.h>
#include <device_launch_parameters.h>
#include <cub/cub.cuh>
#include <cub/device/device_radix_sort.cuh>
#define GLM_FORCE_CUDA
#include <glm/glm.hpp>
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

namespace std {
struct gsref_null_t {};
static gsref_null_t gsref_null_out;
template <class T>
inline gsref_null_t& operator<<(gsref_null_t& s, const T&) { return s; }
inline gsref_null_t& operator<<(gsref_null_t& s, std::ostream& (*)(std::ostream&)) { return s; }
struct gsref_null_ofs {
	gsref_null_ofs(const char*) {}
	void close() {}
};
template <class T>
inline gsref_null_ofs& operator<<(gsref_null_ofs& s, const T&) { return s; }
}  // namespace std

static inline cudaError_t gsref_memcpy(void* dst, const void* src, size_t n, cudaMemcpyKind kind)
{
	if (n != sizeof(int))
		return cudaSuccess;  // debug dump copies: skipped
	return cudaMemcpy(dst, src, n, kind);
}

#define printf(...) ((void)0)
#define cout gsref_null_out
#define ofstream gsref_null_ofs
#define cudaMemcpy gsref_memcpy
