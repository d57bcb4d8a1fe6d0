// Shared definitions for the gnnseg kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gnnseg.h"

namespace gnnseg {

// Offsets (in floats) of the packed weight blob.  Everything is stored transposed
// ([k][out]) so a thread reads 4 consecutive outputs with one 16-byte shared-memory load,
// and every D-wide input block [hidden | X] is padded to D4 = h+4 columns (zero weights on
// the padding) so rows of HX are 16-byte aligned.  Column map: SURVEY.md §3.1 /
// gnn/model.py:73,120,146.
template <int H>
struct Blob {
    static constexpr int D4  = H + 4;
    static constexpr int WIN = 0;                    // [4][H]      input_network.0.weight^T
    static constexpr int BIN = WIN + 4 * H;          // [H]
    static constexpr int W1  = BIN + H;              // [D4][2H]    edge layer 0: [src part | dst part]
    static constexpr int B1  = W1 + D4 * 2 * H;      // [H]
    static constexpr int W2  = B1 + H;               // [H]         edge layer 2
    static constexpr int B2  = W2 + H;               // [4]         b2, 0, 0, 0
    static constexpr int W3  = B2 + 4;               // [3*D4][H]   node layer 0: [mi | mo | self]
    static constexpr int B3  = W3 + 3 * D4 * H;      // [H]
    static constexpr int W4  = B3 + H;               // [H][H]      node layer 2
    static constexpr int B4  = W4 + H * H;           // [H]
    static constexpr int TOTAL = B4 + H;
};

__host__ __device__ inline int blob_total(int h) {
    const int d4 = h + 4;
    return 4 * h + h + d4 * 2 * h + h + h + 4 + 3 * d4 * h + h + h * h + h;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 lds4(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st4(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

// Row stride (floats) for a [rows][k] shared tile read as float4 by up to 8 consecutive
// rows at once: stride/4 odd makes those 8 float4 land on distinct banks.
__host__ __device__ constexpr int tile_stride(int k) { return ((k / 4) & 1) ? k : k + 4; }

}  // namespace gnnseg
