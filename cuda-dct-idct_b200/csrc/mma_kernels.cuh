// mma_kernels.cuh -- the TENSOR-CORE arm of the exact (dense-T) variant: BASELINE configs[3]
// "Exact DCT (cublasDCTv2-equivalent dense 8x8 contraction) on 16384x16384, CUDA-core vs tensor-core
// path".  north_star: "Only the exact (dense) DCT variant may use tensor cores, as a batched 8x8
// contraction, and only if ncu shows it beats the CUDA-core path."  This is that batched 8x8
// contraction, built so that ncu can answer the question (profiles/r02_ncu_k_mma_*.txt; DESIGN.md
// section 6b); it is opt-in (b200dct_dense_mode B200DCT_DENSE_MMA), never chosen by AUTO.
// What it replaces in the reference: the two N x N x N cublasSgemm calls on a block-diagonal T per
// direction (main_cublass_2.cu:228-235,288-295) and the per-block Sgemm pairs of main_cublass.cu:234-241.
//
// One warp owns one 8x8 block at a time as an mma.sync.m16n8k8 (TF32 inputs, FP32 accumulate)
// fragment: lane (g = lane/4, q = lane%4) holds the two pixels (rows 2q and 2q+1, column g).
// All four passes  T.X,  (T.X).T^T,  T^T.D,  (T^T.D).T  run back to back IN REGISTERS: the
// accumulator fragment of one pass is, element for element, the B fragment of the next (the
// contraction index is carried by the free permutation of the k slots, which is folded into the
// constant A fragment), so there is no shuffle, no shared memory and no barrier anywhere.
// FP32 fidelity on TF32 hardware: the constant A operand is the 16 x 8 stack [T_hi ; T_lo]
// (T = T_hi + T_lo, both TF32), and every data operand is issued twice (x_hi, x - x_hi), so one
// pair of MMAs accumulates all four cross terms in FP32; rows 0-7 and 8-15 of the accumulator are
// then added.  Error per inner product ~2^-22 relative: the same order as any FP32 reassociation
// (cuBLAS against the ordered chain: 8-54 of 65536 coefficients differ by one step).
// The quantiser and the +-128 shifts are the CUDA-core ones (two coefficients per lane, whose table
// entries are fixed per lane and live in registers).
#pragma once

#include "dct_kernels.cuh"

namespace b200dct {

struct MmaParams {
    const float *in;
    float *out;
    float *coef;          // optional f32 coefficient plane
    size_t in_pitch, out_pitch, coef_pitch; // bytes
    int bx, by;
    CommonParams cp;
};

__device__ __forceinline__ uint32_t tf32_rna(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// one pass: out = L . in with in = (v0, v1) the lane's B fragment (k slots q and q+4) and A = [L_hi ; L_lo]
__device__ __forceinline__ void mma_pass(const uint32_t (&a)[4], float v0, float v1, float &o0, float &o1)
{
    const uint32_t h0 = tf32_rna(v0), h1 = tf32_rna(v1);
    const float l0 = v0 - __uint_as_float(h0), l1 = v1 - __uint_as_float(h1); // exact; the MMA keeps its top 11 bits
    float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    mma_tf32(d, a, h0, h1);
    mma_tf32(d, a, __float_as_uint(l0), __float_as_uint(l1));
    o0 = d[0] + d[2]; // rows m (hi part of L) + rows m+8 (lo part of L)
    o1 = d[1] + d[3];
}

template <bool FASTDIV>
__global__ void __launch_bounds__(128, 8) k_mma(const __grid_constant__ MmaParams P)
{
    const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
    const long long by = (long long)blockIdx.x * 4 + threadIdx.y;
    const int bx0 = blockIdx.y * 32;
    if (by >= P.by || bx0 >= P.bx) return; // warp-uniform
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // constant A fragments, k slot q <-> index 2q, slot q+4 <-> index 2q+1:
    //   fwd: rows = T[m][.]   (passes 1, 2)      inv: rows = T^T[m][.] = T[.][m]   (passes 3, 4)
    uint32_t afwd[4], ainv[4];
    {
        const float f0 = P.cp.t.t[g * 8 + 2 * q], f1 = P.cp.t.t[g * 8 + 2 * q + 1];
        const float i0 = P.cp.t.tt[g * 8 + 2 * q], i1 = P.cp.t.tt[g * 8 + 2 * q + 1];
        afwd[0] = tf32_rna(f0); afwd[1] = tf32_rna(f0 - __uint_as_float(afwd[0]));
        afwd[2] = tf32_rna(f1); afwd[3] = tf32_rna(f1 - __uint_as_float(afwd[2]));
        ainv[0] = tf32_rna(i0); ainv[1] = tf32_rna(i0 - __uint_as_float(ainv[0]));
        ainv[2] = tf32_rna(i1); ainv[3] = tf32_rna(i1 - __uint_as_float(ainv[2]));
    }
    // after pass 2 the lane holds Y[2q][g] and Y[2q+1][g]: its two quantiser entries never change
    const int k0 = (2 * q) * 8 + g, k1 = (2 * q + 1) * 8 + g;
    const float d0 = P.cp.q.d[k0], d1 = P.cp.q.d[k1], r0 = P.cp.q.rcp[k0], r1 = P.cp.q.rcp[k1];
    const uint32_t m0 = P.cp.q.keep[k0], m1 = P.cp.q.keep[k1];
    auto quant = [&](float y, float d, float r, uint32_t keep) {
        float qv;
        if constexpr (FASTDIV) {
            const float q0 = y * r;
            qv = __fmaf_rn(__fmaf_rn(q0, -d, y), r, q0); // correctly rounded y / d (dct_core.cuh)
        } else {
            qv = __fdiv_rn(y, d);
        }
        return __uint_as_float(__float_as_uint(roundf(qv)) & keep); // divide_matrices, utils_kernels.cu:42
    };

    const int nb = P.bx - bx0 < 32 ? P.bx - bx0 : 32;
    const size_t row_a = ((size_t)by * 8 + 2 * q), col = (size_t)bx0 * 8 + g;
    const char *src0 = (const char *)P.in + row_a * P.in_pitch + col * 4;
    char *dst0 = (char *)P.out + row_a * P.out_pitch + col * 4;
    char *cf0 = P.coef ? (char *)P.coef + row_a * P.coef_pitch + col * 4 : nullptr;
    constexpr int U = 8; // independent blocks per batch; the next batch's loads are in flight while this one computes
    float x0[U], x1[U], n0[U], n1[U];
    auto load_batch = [&](int b, float (&a0)[U], float (&a1)[U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = b + u < nb;
            a0[u] = ok ? __ldg(reinterpret_cast<const float *>(src0 + (size_t)(b + u) * 32)) : 0.0f;
            a1[u] = ok ? __ldg(reinterpret_cast<const float *>(src0 + P.in_pitch + (size_t)(b + u) * 32)) : 0.0f;
        }
    };
    load_batch(0, x0, x1);
    for (int b = 0; b < nb; b += U) {
        if (b + U < nb) load_batch(b + U, n0, n1);
#pragma unroll
        for (int u = 0; u < U; u++) {
            float v0 = x0[u] - 128.0f, v1 = x1[u] - 128.0f; // sub_matrix_scalar, utils_kernels.cu:16
            mma_pass(afwd, v0, v1, v0, v1);                  // M = T.X          lane: M[g][2q], M[g][2q+1]
            mma_pass(afwd, v0, v1, v0, v1);                  // Y^T = T.M^T      lane: Y[2q][g], Y[2q+1][g]
            v0 = quant(v0, d0, r0, m0);
            v1 = quant(v1, d1, r1, m1);
            if (cf0 && b + u < nb) {
                *reinterpret_cast<float *>(cf0 + (size_t)(b + u) * 32) = v0;
                *reinterpret_cast<float *>(cf0 + P.coef_pitch + (size_t)(b + u) * 32) = v1;
            }
            v0 *= d0;                                        // multiply_matrices, utils_kernels.cu:55
            v1 *= d1;
            mma_pass(ainv, v0, v1, v0, v1);                  // M2 = T^T.D       lane: M2[g][2q], M2[g][2q+1]
            mma_pass(ainv, v0, v1, v0, v1);                  // R^T = T^T.M2^T   lane: R[2q][g], R[2q+1][g]
            if (b + u < nb) {
                *reinterpret_cast<float *>(dst0 + (size_t)(b + u) * 32) = v0 + 128.0f; // add_matrix_scalar, utils_kernels.cu:29
                *reinterpret_cast<float *>(dst0 + P.out_pitch + (size_t)(b + u) * 32) = v1 + 128.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            x0[u] = n0[u];
            x1[u] = n1[u];
        }
    }
}

} // namespace b200dct
