// Per-optimizer-step kernels (K3/K4) and the evaluation kernel (K5).
//
// Shared-evaluation path (cell size == grid spacing): the network is evaluated ONCE per node of
// the rank-local lattice (fwd_nodes); the rows (7-point stencil on U) and the adjoint stencil
// (d loss/d U per node) of the face table are ONE kernel fed by 3-D TMA boxes (nbm_stencil_tma.cuh);
// crossed nodes get their far-side value from the 27-point regression table and irregular rows
// their corrections in a chain of list kernels that runs on a side stream BESIDE that kernel
// (extrap -> irregular_fb -> extrap_bwd, merged into G by merge_lists); one fused forward-recompute
// + backward kernel (node_grad: 12 warps/SM, pair-split outer-product accumulators) turns G into
// per-CTA partial sums of d loss/d theta held in registers; finalize_step (or, on several GPUs,
// reduce_allreduce_finalize over NVLink peer memory) sums them and runs the optax chain.
// Optional: the learned preconditioner kernels between separate residual and adjoint kernels.
// General path (any cell size): the network kernels over 7 displaced lattices - 4 shared ones at
// zoom level 1 -, a pointwise rows kernel, and the 27-cubes of crossed sites as 3x3x3 mini-lattices.
//
// The network parameters live in __constant__ memory so that every FFMA takes its weight as a
// constant-bank operand (no load, no register).
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "nbm_common.cuh"

namespace nbm {

#define NBM_MAXP 1024
// [0, MAXP): the parameters; [MAXP, 2*MAXP): the same with the hidden layers' W and b multiplied
// by 2*log2(e), so that the forward pass feeds ex2 directly (tanh(s) = 1 - 2/(2^(2 log2e s) + 1));
// [2*MAXP, 3*MAXP): the parameters with every hidden HxH matrix TRANSPOSED, so that the backward
// product W delta also runs as 5 independent fp32x2 chains with adjacent constant-bank pairs.
__constant__ __align__(16) float c_P[3 * NBM_MAXP];
__device__ __align__(16) float g_stage[3 * NBM_MAXP];
// the learned preconditioner's parameters (nn/preconditioner.py), uploaded before its kernels run
__constant__ __align__(16) float c_PC[512];
constexpr float kTwoLog2e = 2.8853900817779268f;

constexpr int kThreads = 256;

// tanh(x) = 1 - 2/(exp(2x)+1): 2 MUFU (ex2, rcp) + 3 FP32 ops, abs error ~1.5e-7 (tanh.approx's
// 2^-11 is not enough for the 1e-5 residual tolerance).  Saturates correctly at +-inf.
__device__ __forceinline__ float tanh_nbm(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

// hk.Linear stack (MLP.py:111-116, 128-139): y = x W + b, W (in,out) row-major, tanh on hidden layers.
// Flat layout per head: W1(3xH) b1(H) [W(HxH) b(H)]*(L-1) Wout(H) bout(1).
template <int L, int H>
struct Mlp {
    static constexpr int NP = 3 * H + H + (L - 1) * (H * H + H) + H + 1;

    template <int OFF>
    __device__ __forceinline__ static float forward(float x, float y, float z, float (&a)[L][H]) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
            float s = c_P[OFF + 3 * H + j];
            s = fmaf(x, c_P[OFF + j], s);
            s = fmaf(y, c_P[OFF + H + j], s);
            s = fmaf(z, c_P[OFF + 2 * H + j], s);
            a[0][j] = tanh_nbm(s);
        }
#pragma unroll
        for (int l = 1; l < L; ++l) {
            const int o = OFF + 4 * H + (l - 1) * (H * H + H);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                float s = c_P[o + H * H + j];
#pragma unroll
                for (int i = 0; i < H; ++i) s = fmaf(a[l - 1][i], c_P[o + i * H + j], s);
                a[l][j] = tanh_nbm(s);
            }
        }
        const int o = OFF + 4 * H + (L - 1) * (H * H + H);
        float out = c_P[o + H];
#pragma unroll
        for (int i = 0; i < H; ++i) out = fmaf(a[L - 1][i], c_P[o + i], out);
        return out;
    }

    // acc[OFF..OFF+NP) += g * d u / d theta
    // acc[k] accumulates parameter (BASE + k); OFF is the head's offset in the constant bank
    template <int OFF, int NTOT, int BASE>
    __device__ __forceinline__ static void backward(float x, float y, float z, const float (&a)[L][H], float g,
                                                    float (&acc_)[NTOT]) {
        const int oo = OFF + 4 * H + (L - 1) * (H * H + H);
        float* acc = acc_ - BASE;  // compile-time indices after unrolling: stays in registers
        float d[H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            acc[oo + i] = fmaf(a[L - 1][i], g, acc[oo + i]);
            d[i] = c_P[oo + i] * g * fmaf(-a[L - 1][i], a[L - 1][i], 1.0f);
        }
        acc[oo + H] += g;
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            const int o = OFF + 4 * H + (l - 1) * (H * H + H);
            float dn[H];
#pragma unroll
            for (int i = 0; i < H; ++i) dn[i] = 0.0f;
#pragma unroll
            for (int j = 0; j < H; ++j) acc[o + H * H + j] += d[j];
#pragma unroll
            for (int i = 0; i < H; ++i) {
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    acc[o + i * H + j] = fmaf(a[l - 1][i], d[j], acc[o + i * H + j]);
                    dn[i] = fmaf(c_P[o + i * H + j], d[j], dn[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < H; ++i) d[i] = dn[i] * fmaf(-a[l - 1][i], a[l - 1][i], 1.0f);
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
            acc[OFF + 3 * H + j] += d[j];
            acc[OFF + j] = fmaf(x, d[j], acc[OFF + j]);
            acc[OFF + H + j] = fmaf(y, d[j], acc[OFF + H + j]);
            acc[OFF + 2 * H + j] = fmaf(z, d[j], acc[OFF + 2 * H + j]);
        }
    }

    // gradient of the head w.r.t. the position (K5: jax.grad of solution_at_point_fn, trainer.py:953-957)
    template <int OFF>
    __device__ __forceinline__ static void input_grad(const float (&a)[L][H], float (&gx)[3]) {
        const int oo = OFF + 4 * H + (L - 1) * (H * H + H);
        float d[H];
#pragma unroll
        for (int i = 0; i < H; ++i) d[i] = c_P[oo + i] * fmaf(-a[L - 1][i], a[L - 1][i], 1.0f);
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            const int o = OFF + 4 * H + (l - 1) * (H * H + H);
            float dn[H];
#pragma unroll
            for (int i = 0; i < H; ++i) {
                dn[i] = 0.0f;
#pragma unroll
                for (int j = 0; j < H; ++j) dn[i] = fmaf(c_P[o + i * H + j], d[j], dn[i]);
            }
#pragma unroll
            for (int i = 0; i < H; ++i) d[i] = dn[i] * fmaf(-a[l - 1][i], a[l - 1][i], 1.0f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < H; ++j) s = fmaf(c_P[OFF + c * H + j], d[j], s);
            gx[c] = s;
        }
    }
};


// ---------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100+).  Measured on B200 (tools/microbench.cu):
// FFMA2 retires the same 128 lane-FMAs/clk/SM as FFMA but takes half the issue slots, which lets the
// MUFU (ex2, rcp) and load instructions co-issue for free.  Pairs run along the FEATURE index
// (j, j+1): weights W[i][j], W[i][j+1] are adjacent in the (in,out) row-major layout and reach the
// FFMA2 straight from the constant bank through a uniform register pair.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo32(u64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi32(u64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 neg2(u64 a) { return a ^ 0x8000000080000000ull; }
__device__ __forceinline__ u64 cpair(int idx) { return *reinterpret_cast<const u64*>(&c_P[idx]); }

// two tanh for the price of 3 MUFU: t = 1 - 2/(e+1) with ONE reciprocal of (e0+1)(e1+1).
// Input is the PRE-SCALED pre-activation s' = 2 log2(e) s.  Clamped at 63 so the product stays finite
// (tanh is 1 to the last bit far below that).
__device__ __forceinline__ u64 tanh2_prescaled(u64 sp) {
    float s0 = fminf(lo32(sp), 63.0f), s1 = fminf(hi32(sp), 63.0f);
    float e0, e1, m;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
    // A = (a1, -a0/2) with a = e + 1;  m = 1 / (lo hi) = -2 / (a0 a1);  t = 1 + (m, -2m) A = (1 - 2/a0, 1 - 2/a1)
    const u64 A = ffma2(pk(e1, e0), pk(1.0f, -0.5f), pk(1.0f, -0.5f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(lo32(A) * hi32(A)));
    return ffma2(pk(m, -2.0f * m), A, pk(1.0f, 1.0f));
}

// FOUR tanh for the price of 5 MUFU (4 ex2 + ONE reciprocal of the product of the four (e + 1)): the forward kernel is
// MUFU-bound (XU pipe 76 % busy with the pair version), this trades 5 MUFU per two pairs for 3 more multiplies.  Clamped
// at 31: (2^31 + 1)^4 stays finite and 1 - 2/(2^31 + 1) already rounds to 1.  1/a_i = (product of the other three) / P
// costs two roundings more than the pair version: abs error ~4e-7.
__device__ __forceinline__ void tanh4_prescaled(u64 sp0, u64 sp1, u64& t0, u64& t1) {
    const float s0 = fminf(lo32(sp0), 31.0f), s1 = fminf(hi32(sp0), 31.0f), s2 = fminf(lo32(sp1), 31.0f),
                s3 = fminf(hi32(sp1), 31.0f);
    float e0, e1, e2, e3, m;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(s2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(s3));
    const u64 one2 = pk(1.0f, 1.0f);
    const u64 A = fadd2(pk(e0, e1), one2), B = fadd2(pk(e2, e3), one2);   // (a0, a1), (a2, a3)
    const float pA = lo32(A) * hi32(A), pB = lo32(B) * hi32(B);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(pA * pB));
    const float rA = -2.0f * (m * pB), rB = -2.0f * (m * pA);               // -2 / (a0 a1), -2 / (a2 a3)
    // t_i = 1 - 2 / a_i = 1 + r (the partner of a_i inside its pair)
    t0 = ffma2(pk(rA, rA), pk(hi32(A), lo32(A)), one2);
    t1 = ffma2(pk(rB, rB), pk(hi32(B), lo32(B)), one2);
}

// Variant for kernels with MUFU head-room (node_grad: XU pipe ~30 % busy): one reciprocal per element instead of a
// shared one.  4 MUFU per pair instead of 3, but 6 instructions instead of 9 and no clamp: rcp(inf) = 0 gives t = 1.
__device__ __forceinline__ u64 tanh2_prescaled_4mufu(u64 sp) {
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(lo32(sp)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(hi32(sp)));
    const u64 a = fadd2(pk(e0, e1), pk(1.0f, 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(lo32(a)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(hi32(a)));
    return ffma2(pk(r0, r1), pk(-2.0f, -2.0f), pk(1.0f, 1.0f));
}

// Packed head: H even, OFF even.  Accumulators are pairs: accp[k] = (grad[2k], grad[2k+1]).
template <int L, int H>
struct MlpP {
    static_assert(H % 2 == 0, "packed head needs an even width");
    static constexpr int NP = 3 * H + H + (L - 1) * (H * H + H) + H + 1;
    static constexpr int NPAIR = (NP + 1) / 2;
    static constexpr int HP = H / 2;

    template <int OFF>
    __device__ __forceinline__ static float forward(float x, float y, float z, u64 (&a)[L][HP]) {
        constexpr int S = NBM_MAXP + OFF;  // pre-scaled copy
        const u64 x2 = pk(x, x), y2 = pk(y, y), z2 = pk(z, z);
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            u64 s = cpair(S + 3 * H + 2 * j);
            s = ffma2(x2, cpair(S + 2 * j), s);
            s = ffma2(y2, cpair(S + H + 2 * j), s);
            s = ffma2(z2, cpair(S + 2 * H + 2 * j), s);
            a[0][j] = tanh2_prescaled(s);
        }
#pragma unroll
        for (int l = 1; l < L; ++l) {
            const int o = S + 4 * H + (l - 1) * (H * H + H);
            u64 acc[HP];
#pragma unroll
            for (int j = 0; j < HP; ++j) acc[j] = cpair(o + H * H + 2 * j);
#pragma unroll
            for (int i = 0; i < H; ++i) {
                float ai = (i & 1) ? hi32(a[l - 1][i / 2]) : lo32(a[l - 1][i / 2]);
                u64 ai2 = pk(ai, ai);
#pragma unroll
                for (int j = 0; j < HP; ++j) acc[j] = ffma2(ai2, cpair(o + i * H + 2 * j), acc[j]);
            }
#pragma unroll
            for (int j = 0; j < HP; ++j) a[l][j] = tanh2_prescaled(acc[j]);
        }
        const int oo = OFF + 4 * H + (L - 1) * (H * H + H);  // output layer: unscaled copy
        u64 o2 = pk(c_P[oo + H], 0.0f);
#pragma unroll
        for (int j = 0; j < HP; ++j) o2 = ffma2(a[L - 1][j], cpair(oo + 2 * j), o2);
        return lo32(o2) + hi32(o2);
    }

    // Activation stash: the last hidden layer of every plus-side node, written by the forward kernel as HPQ planes of
    // 16-byte chunks (two feature pairs each; a warp writes 512 contiguous bytes per plane) and read back by the
    // gradient kernel, which then recomputes only the layers below it.
    static constexpr int HPQ = (HP + 1) / 2;
    __device__ __forceinline__ static void stash_store(float4* __restrict__ hp, int64_t hstride, const u64 (&a)[HP]) {
#pragma unroll
        for (int q = 0; q < HPQ; ++q) {
            const u64 lo = a[2 * q], hi = (2 * q + 1 < HP) ? a[2 * q + 1] : 0ull;
            __stcs(hp + q * hstride, make_float4(lo32(lo), hi32(lo), lo32(hi), hi32(hi)));
        }
    }

    // forward of NB nodes that share (y, z): every constant-bank pair is fetched once and used NB times, and the
    // NB dependency chains interleave (more ILP for the ex2/rcp latency)
    // the part of the first layer that does not depend on x: b1 + y W1[1] + z W1[2] (pre-scaled); a thread that
    // marches along x with fixed (y, z) computes it once per task
    template <int OFF>
    __device__ __forceinline__ static void first_layer_yz(float y, float z, u64 (&yz)[HP]) {
        constexpr int S = NBM_MAXP + OFF;
        const u64 y2 = pk(y, y), z2 = pk(z, z);
#pragma unroll
        for (int j = 0; j < HP; ++j)
            yz[j] = ffma2(z2, cpair(S + 2 * H + 2 * j), ffma2(y2, cpair(S + H + 2 * j), cpair(S + 3 * H + 2 * j)));
    }

    template <int OFF, int NB>
    __device__ __forceinline__ static void forward_many(const float (&x)[NB], const u64 (&yz)[HP], float (&out)[NB],
                                                        float4* __restrict__ hp = nullptr, int64_t hstride = 0,
                                                        int64_t hnode = 0) {
        constexpr int S = NBM_MAXP + OFF;
        u64 a[NB][HP], b[NB][HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            const u64 wx = cpair(S + 2 * j);
            if (NB == 2) {   // the two nodes' pairs share one reciprocal
                tanh4_prescaled(ffma2(pk(x[0], x[0]), wx, yz[j]), ffma2(pk(x[NB - 1], x[NB - 1]), wx, yz[j]), a[0][j], a[NB - 1][j]);
            } else {
#pragma unroll
                for (int n = 0; n < NB; ++n) a[n][j] = tanh2_prescaled(ffma2(pk(x[n], x[n]), wx, yz[j]));
            }
        }
#pragma unroll
        for (int l = 1; l < L; ++l) {
            const int o = S + 4 * H + (l - 1) * (H * H + H);
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int j = 0; j < HP; ++j) b[n][j] = cpair(o + H * H + 2 * j);
#pragma unroll
            for (int i = 0; i < H; ++i) {
#pragma unroll
                for (int j = 0; j < HP; ++j) {
                    const u64 w = cpair(o + i * H + 2 * j);
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        float ai = (i & 1) ? hi32(a[n][i / 2]) : lo32(a[n][i / 2]);
                        b[n][j] = ffma2(pk(ai, ai), w, b[n][j]);
                    }
                }
            }
            if (NB == 2) {
#pragma unroll
                for (int j = 0; j < HP; ++j) tanh4_prescaled(b[0][j], b[NB - 1][j], a[0][j], a[NB - 1][j]);
            } else {
#pragma unroll
                for (int n = 0; n < NB; ++n)
#pragma unroll
                    for (int j = 0; j < HP; ++j) a[n][j] = tanh2_prescaled(b[n][j]);
            }
        }
        const int oo = OFF + 4 * H + (L - 1) * (H * H + H);
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            u64 o2 = pk(c_P[oo + H], 0.0f);
#pragma unroll
            for (int j = 0; j < HP; ++j) o2 = ffma2(a[n][j], cpair(oo + 2 * j), o2);
            out[n] = lo32(o2) + hi32(o2);
            if (hp) stash_store(hp + n * hnode, hstride, a[n]);   // node n of the batch sits hnode chunks further
        }
    }

    template <int OFF>
    __device__ __forceinline__ static void backward(float x, float y, float z, const u64 (&a)[L][HP], float g,
                                                    u64 (&acc)[NPAIR]) {
        static_assert(OFF == 0, "packed accumulators are indexed from the head's own origin");
        const int oo = 4 * H + (L - 1) * (H * H + H);
        const u64 mone = pk(-1.0f, -1.0f);
        const u64 g2 = pk(g, g), g2n = pk(-g, -g);
        u64 d[HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            acc[(oo + 2 * j) / 2] = ffma2(a[L - 1][j], g2, acc[(oo + 2 * j) / 2]);
            u64 omn = ffma2(a[L - 1][j], a[L - 1][j], mone);          // a^2 - 1
            d[j] = fmul2(fmul2(cpair(oo + 2 * j), g2n), omn);         // W g (1 - a^2)
        }
        acc[(oo + H) / 2] = fadd2(acc[(oo + H) / 2], pk(g, 0.0f));
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            const int o = 4 * H + (l - 1) * (H * H + H);
#pragma unroll
            for (int j = 0; j < HP; ++j) acc[(o + H * H + 2 * j) / 2] = fadd2(acc[(o + H * H + 2 * j) / 2], d[j]);
            // outer product  dW[i][j] += a_i delta_j  (pairs along j)
#pragma unroll
            for (int i = 0; i < H; ++i) {
                float ai = (i & 1) ? hi32(a[l - 1][i / 2]) : lo32(a[l - 1][i / 2]);
                u64 ai2 = pk(ai, ai);
#pragma unroll
                for (int j = 0; j < HP; ++j)
                    acc[(o + i * H + 2 * j) / 2] = ffma2(ai2, d[j], acc[(o + i * H + 2 * j) / 2]);
            }
            // delta_prev = (W delta) (1 - a^2): pairs along i through the transposed copy -> HP independent chains
            u64 dn[HP];
#pragma unroll
            for (int ip = 0; ip < HP; ++ip) dn[ip] = 0ull;
#pragma unroll
            for (int j = 0; j < H; ++j) {
                float dj = (j & 1) ? hi32(d[j / 2]) : lo32(d[j / 2]);
                u64 dj2 = pk(dj, dj);
#pragma unroll
                for (int ip = 0; ip < HP; ++ip) dn[ip] = ffma2(cpair(2 * NBM_MAXP + o + j * H + 2 * ip), dj2, dn[ip]);
            }
#pragma unroll
            for (int ip = 0; ip < HP; ++ip) {   // dn = -(W delta) from the negated transposed copy
                u64 omn = ffma2(a[l - 1][ip], a[l - 1][ip], mone);
                dn[ip] = fmul2(dn[ip], omn);
            }
#pragma unroll
            for (int j = 0; j < HP; ++j) d[j] = dn[j];
        }
        const u64 x2 = pk(x, x), y2 = pk(y, y), z2 = pk(z, z);
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            acc[(3 * H + 2 * j) / 2] = fadd2(acc[(3 * H + 2 * j) / 2], d[j]);
            acc[(2 * j) / 2] = ffma2(x2, d[j], acc[(2 * j) / 2]);
            acc[(H + 2 * j) / 2] = ffma2(y2, d[j], acc[(H + 2 * j) / 2]);
            acc[(2 * H + 2 * j) / 2] = ffma2(z2, d[j], acc[(2 * H + 2 * j) / 2]);
        }
    }
    // -----------------------------------------------------------------------------------------
    // Pair-split backward (node_grad): the two lanes of a pair (lane ^ 1) walk neighbouring nodes; the
    // H x H outer-product accumulators of every hidden layer are split between them by ROW PARITY (lane
    // parity b owns rows i = 2k + b) and each lane accumulates its rows for BOTH nodes, receiving the
    // partner's delta (H values) and the partner's activations of its rows (H/2 values) by shuffle.
    // That halves the largest accumulator block (100 -> 50 registers for H = 10).  The first layer is
    // hoisted along the x march: (y, z) are fixed per task, so only sum(delta1) and sum(x delta1) are
    // kept per node; the y, z rows and the bias follow from sum(delta1) at the end of the task.
    // -----------------------------------------------------------------------------------------
    static constexpr int LH = (L > 1) ? (L - 1) : 1;
    struct AccS {
        u64 w3[HP];
        float b3;
        u64 w2[LH][HP][HP];   // [layer][k: row 2k + parity][pair of columns]
        u64 b2[LH][HP];
        u64 w1x[HP];
        u64 t[HP];            // task-local sum of delta1
        __device__ __forceinline__ void zero() {
            b3 = 0.0f;
#pragma unroll
            for (int j = 0; j < HP; ++j) { w3[j] = 0ull; w1x[j] = 0ull; t[j] = 0ull; }
#pragma unroll
            for (int l = 0; l < LH; ++l)
#pragma unroll
                for (int k = 0; k < HP; ++k) {
                    b2[l][k] = 0ull;
#pragma unroll
                    for (int j = 0; j < HP; ++j) w2[l][k][j] = 0ull;
                }
        }
    };

    __device__ __forceinline__ static u64 shfl1(u64 v) {
        float a = __shfl_xor_sync(0xffffffffu, lo32(v), 1), b = __shfl_xor_sync(0xffffffffu, hi32(v), 1);
        return pk(a, b);
    }

    // forward recompute (first layer from the hoisted yz part) + backward with g = d loss / d u(node).
    // Must be called by all 32 lanes (g = 0 for lanes that have nothing to add).
    // STASH: the last hidden layer `aLast` comes from the forward kernel's stash instead of being recomputed.
    template <bool STASH = false>
    __device__ __forceinline__ static void grad_split(float x, const u64 (&yz)[HP], float g, AccS& A, bool par,
                                                      const u64* aLast = nullptr) {
        constexpr int S = NBM_MAXP;
        u64 a[L][HP];
        const u64 x2 = pk(x, x);
        if (STASH) {
#pragma unroll
            for (int j = 0; j < HP; ++j) a[L - 1][j] = aLast[j];
        }
        if (!STASH || L > 1) {
#pragma unroll
            for (int j = 0; j < HP; ++j) a[0][j] = tanh2_prescaled_4mufu(ffma2(x2, cpair(S + 2 * j), yz[j]));
        }
#pragma unroll
        for (int l = 1; l < (STASH ? L - 1 : L); ++l) {
            const int o = S + 4 * H + (l - 1) * (H * H + H);
            u64 acc[HP];
#pragma unroll
            for (int j = 0; j < HP; ++j) acc[j] = cpair(o + H * H + 2 * j);
#pragma unroll
            for (int i = 0; i < H; ++i) {
                float ai = (i & 1) ? hi32(a[l - 1][i / 2]) : lo32(a[l - 1][i / 2]);
                u64 ai2 = pk(ai, ai);
#pragma unroll
                for (int j = 0; j < HP; ++j) acc[j] = ffma2(ai2, cpair(o + i * H + 2 * j), acc[j]);
            }
#pragma unroll
            for (int j = 0; j < HP; ++j) a[l][j] = tanh2_prescaled_4mufu(acc[j]);
        }
        const int oo = 4 * H + (L - 1) * (H * H + H);
        const u64 mone = pk(-1.0f, -1.0f);
        const u64 g2 = pk(g, g), g2n = pk(-g, -g);
        u64 d[HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            A.w3[j] = ffma2(a[L - 1][j], g2, A.w3[j]);
            u64 omn = ffma2(a[L - 1][j], a[L - 1][j], mone);          // a^2 - 1
            d[j] = fmul2(fmul2(cpair(oo + 2 * j), g2n), omn);         // W g (1 - a^2)
        }
        A.b3 += g;
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            const int o = 4 * H + (l - 1) * (H * H + H);
            u64 dP[HP];
#pragma unroll
            for (int j = 0; j < HP; ++j) {
                A.b2[l - 1][j] = fadd2(A.b2[l - 1][j], d[j]);
                dP[j] = shfl1(d[j]);
            }
#pragma unroll
            for (int k = 0; k < HP; ++k) {
                // my row 2k + par of this layer's outer product: my activation and the partner's (selected and
                // exchanged right where they are used: no staging arrays)
                const float lo = lo32(a[l - 1][k]), hi = hi32(a[l - 1][k]);
                const float aown = par ? hi : lo;
                const float apart = __shfl_xor_sync(0xffffffffu, par ? lo : hi, 1);
                const u64 ao = pk(aown, aown), ap = pk(apart, apart);
#pragma unroll
                for (int j = 0; j < HP; ++j)
                    A.w2[l - 1][k][j] = ffma2(ap, dP[j], ffma2(ao, d[j], A.w2[l - 1][k][j]));
            }
            u64 dn[HP];
#pragma unroll
            for (int ip = 0; ip < HP; ++ip) dn[ip] = 0ull;
#pragma unroll
            for (int j = 0; j < H; ++j) {
                float dj = (j & 1) ? hi32(d[j / 2]) : lo32(d[j / 2]);
                u64 dj2 = pk(dj, dj);
#pragma unroll
                for (int ip = 0; ip < HP; ++ip) dn[ip] = ffma2(cpair(2 * NBM_MAXP + o + j * H + 2 * ip), dj2, dn[ip]);
            }
#pragma unroll
            for (int ip = 0; ip < HP; ++ip) {   // dn = -(W delta) from the negated transposed copy
                u64 omn = ffma2(a[l - 1][ip], a[l - 1][ip], mone);
                d[ip] = fmul2(dn[ip], omn);
            }
        }
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            A.t[j] = fadd2(A.t[j], d[j]);
            A.w1x[j] = ffma2(x2, d[j], A.w1x[j]);
        }
    }

    // the value this lane contributes to flat gradient entry i of the head (compile-time i); `hs` are the
    // thread's 3H hoisted sums in shared memory: [0,H) bias, [H,2H) y row, [2H,3H) z row (stride `hstride`)
    __device__ __forceinline__ static float split_get(const AccS& A, int i, bool par, const float* hs, int hstride) {
        const int oo = 4 * H + (L - 1) * (H * H + H);
        if (i < H) return (i & 1) ? hi32(A.w1x[i / 2]) : lo32(A.w1x[i / 2]);
        if (i < 3 * H) return hs[i * hstride];
        if (i < 4 * H) return hs[(i - 3 * H) * hstride];
        if (i < oo) {
            const int l = (i - 4 * H) / (H * H + H), r = (i - 4 * H) - l * (H * H + H);
            if (r >= H * H) {
                const int j = r - H * H;
                return (j & 1) ? hi32(A.b2[l][j / 2]) : lo32(A.b2[l][j / 2]);
            }
            const int row = r / H, j = r - row * H;
            const u64 v = A.w2[l][row / 2][j / 2];
            const float f = (j & 1) ? hi32(v) : lo32(v);
            return (par == ((row & 1) != 0)) ? f : 0.0f;
        }
        if (i < oo + H) return ((i - oo) & 1) ? hi32(A.w3[(i - oo) / 2]) : lo32(A.w3[(i - oo) / 2]);
        return A.b3;
    }
};

template <int LP, int HP, int LM, int HM>
struct Net {
    using P = MlpP<LP, HP>;
    using M = Mlp<LM, HM>;
    static constexpr int NP = P::NP + M::NP;
    static constexpr int HPW = HP, LMD = LM, HMW = HM;

    // per-thread gradient accumulators: p-head in fp32x2 pairs, m-head scalar
    struct Acc {
        u64 p[P::NPAIR];
        float m[M::NP];
        __device__ __forceinline__ void zero() {
#pragma unroll
            for (int i = 0; i < P::NPAIR; ++i) p[i] = 0ull;
#pragma unroll
            for (int i = 0; i < M::NP; ++i) m[i] = 0.0f;
        }
        // gradient entry i in the flat C-ABI order
        __device__ __forceinline__ float get(int i) const {
            return i < P::NP ? ((i & 1) ? hi32(p[i / 2]) : lo32(p[i / 2])) : m[i - P::NP];
        }
    };

    // DoubleMLP.__call__ (MLP.py:98): phi >= 0 ? mlp_p : mlp_m
    __device__ __forceinline__ static float eval(bool plus, float x, float y, float z) {
        if (plus) {
            u64 a[LP][HP / 2];
            return P::template forward<0>(x, y, z, a);
        } else {
            float a[LM][HM];
            return M::template forward<P::NP>(x, y, z, a);
        }
    }
    // two nodes (x0, y, z), (x1, y, z) at once; falls back to two single evaluations when the heads differ
    using YZ = u64[HP / 2];
    __device__ __forceinline__ static void first_layer_yz(float y, float z, u64 (&yz)[HP / 2]) {
        P::template first_layer_yz<0>(y, z, yz);
    }
    // `hp`: stash of node 0's last hidden layer (node 1 sits `hnode` 16-byte chunks further), null = not kept
    __device__ __forceinline__ static void eval2(bool plus0, bool plus1, float x0, float x1, float y, float z,
                                                 const u64 (&yz)[HP / 2], float& u0, float& u1, float4* hp = nullptr,
                                                 int64_t hstride = 0, int64_t hnode = 0) {
        if (plus0 && plus1) {
            const float xs[2] = {x0, x1};
            float out[2];
            P::template forward_many<0, 2>(xs, yz, out, hp, hstride, hnode);
            u0 = out[0];
            u1 = out[1];
        } else {
            u0 = eval_stash(plus0, x0, y, z, hp, hstride);
            u1 = eval_stash(plus1, x1, y, z, hp ? hp + hnode : nullptr, hstride);
        }
    }
    __device__ __forceinline__ static float eval_stash(bool plus, float x, float y, float z, float4* hp, int64_t hstride) {
        if (plus) {
            u64 a[LP][HP / 2];
            const float u = P::template forward<0>(x, y, z, a);
            if (hp) P::stash_store(hp, hstride, a[LP - 1]);
            return u;
        }
        float a[LM][HM];
        return M::template forward<P::NP>(x, y, z, a);
    }
    __device__ __forceinline__ static void grad(bool plus, float x, float y, float z, float g, Acc& acc) {
        if (plus) {
            u64 a[LP][HP / 2];
            P::template forward<0>(x, y, z, a);
            P::template backward<0>(x, y, z, a, g, acc.p);
        } else {
            float a[LM][HM];
            M::template forward<P::NP>(x, y, z, a);
            M::template backward<P::NP, M::NP, P::NP>(x, y, z, a, g, acc.m);
        }
    }
};

// nonlinear operator N(u) and N'(u) (discretization.py:369; examples/biomolecules/coefficients.py:126-131)
__device__ __forceinline__ float nl_apply(int kind, float coef, float u) {
    return kind == NBM_NL_SINH ? coef * sinhf(u) : 0.0f;
}
__device__ __forceinline__ float nl_deriv(int kind, float coef, float u) {
    return kind == NBM_NL_SINH ? coef * coshf(u) : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// task decomposition shared by fwd_nodes and node_grad: the (y,z) plane is cut into strips of
// kThreads consecutive cells (z fastest -> coalesced), x into chunks; a CTA walks tasks with a
// grid stride, a thread keeps its (y,z) and marches along x.
// ---------------------------------------------------------------------------------------------
struct Tasks {
    int plane, mblocks, xchunk, nxch, total;
    int split;   // ranges of (strip, x plane) pairs per CTA of the two network kernels (balanced runs), >= 1
};
// ranges per CTA: up to 4 (regions of cheap minus-side nodes spread over more CTAs) while a range keeps >= 64 planes
static int run_split(int64_t plane_iterations, int ctas) {
    static const int forced = getenv("NBM_GRAD_SPLIT") ? atoi(getenv("NBM_GRAD_SPLIT")) : 0;   // timing experiments
    if (forced > 0) return forced;
    const int64_t per_cta = plane_iterations / (ctas > 0 ? ctas : 1);
    return (int)(per_cta >= 256 ? 4 : (per_cta >= 128 ? 2 : 1));
}
__host__ __device__ inline Tasks make_tasks(int ex, int ey, int ez, int xchunk, int threads = kThreads) {
    Tasks t;
    t.plane = ey * ez;
    t.mblocks = (t.plane + threads - 1) / threads;
    t.xchunk = xchunk;
    t.nxch = (ex + xchunk - 1) / xchunk;
    t.total = t.mblocks * t.nxch;
    t.split = 1;
    return t;
}

// What the two network kernels see: `nrep` replicas of a lattice (1 for the shared path; 7 displaced copies of
// the training grid for the general path), each with its own coordinate arrays and its own slice of side/U/G.
struct NodeView {
    const float *xe, *ye, *ze;   // [nrep][ex], [nrep][ey], [nrep][ez]
    int ex, ey, ez;
    int x_begin, x_end;          // planes walked
    int64_t lo, hi;              // only nodes lo <= e < hi (index inside a replica) are touched
    int64_t rep_nodes;           // replica stride of side / U / G
    int nrep;                    // replicas walked (1: shared path, 7: general path); ranges run across replicas
    const uint8_t* side;
    float* U;
    const float* G;
    float* Gz;                   // general path on shared lattices: the forward kernel clears G of the nodes it visits
    const float* R;              // may be null (no loss accumulation)
    float4* Hst;                 // activation stash [HPQ][rep_nodes] 16-byte chunks, or null (recompute)
    float inv_n;
    float* partials;
    int row0;
    int row_stride, loss_col;    // 0: rows are NP + 1 floats with the loss last
};

static NodeView view_of(const nbm_shared_step_t& s) {
    NodeView v;
    v.xe = s.xe; v.ye = s.ye; v.ze = s.ze;
    v.ex = s.ex; v.ey = s.ey; v.ez = s.ez;
    v.x_begin = 0; v.x_end = s.ex;
    v.lo = 0; v.hi = (int64_t)s.ex * s.ey * s.ez;
    v.rep_nodes = v.hi;
    v.nrep = 1;
    v.side = s.side; v.U = s.U; v.G = s.G; v.R = s.R; v.Gz = nullptr;
    v.Hst = reinterpret_cast<float4*>(s.Hst);
    v.inv_n = s.inv_n_points; v.partials = s.partials; v.row0 = 0;
    v.row_stride = 0; v.loss_col = 0;
    return v;
}

// A: U[e] = u(node e)   (evaluate_solution_fn, trainer.py:836-844)
// one thread's part of a task: cell `m` of the flattened (y,z) plane, x planes [x0, x1)
template <class NET, bool GENERAL, bool STASH>
__device__ __forceinline__ void fwd_task(const NodeView& v, const int plane, const float* __restrict__ xe,
                                         const float* __restrict__ ye, const float* __restrict__ ze,
                                         const uint8_t* __restrict__ side, float* __restrict__ U, const int m, const int x0,
                                         const int x1, float* __restrict__ Gz = nullptr) {
    if (m >= plane) return;
    const int iy = m / v.ez, iz = m - iy * v.ez;
    const float y = __ldg(ye + iy), z = __ldg(ze + iz);
    typename NET::YZ yz;
    NET::first_layer_yz(y, z, yz);
    int64_t e = (int64_t)x0 * plane + m;
    // two x planes per iteration (they share y, z and every weight fetch); loads one pair ahead
    float xa_n = __ldg(xe + x0), xb_n = (x0 + 1 < x1) ? __ldg(xe + x0 + 1) : 0.0f;
    uint8_t sa_n = __ldg(side + e), sb_n = (x0 + 1 < x1) ? __ldg(side + e + plane) : (uint8_t)0;
    for (int ix = x0; ix < x1; ix += 2) {
        const float xa = xa_n, xb = xb_n;
        const bool pa = (sa_n & 1) != 0, pb = (sb_n & 1) != 0;
        const int64_t e_cur = e;
        const bool has_b = ix + 1 < x1;
        if (ix + 2 < x1) {
            e += 2 * (int64_t)plane;
            sa_n = __ldg(side + e);
            xa_n = __ldg(xe + ix + 2);
            if (ix + 3 < x1) {
                sb_n = __ldg(side + e + plane);
                xb_n = __ldg(xe + ix + 3);
            }
        }
        float ua, ub;
        float4* hst = STASH ? v.Hst + e_cur : nullptr;   // compile-time null: the default path carries no stash code
        if (has_b) {
            NET::eval2(pa, pb, xa, xb, y, z, yz, ua, ub, hst, v.rep_nodes, plane);
        } else {
            ua = NET::eval_stash(pa, xa, y, z, hst, v.rep_nodes);
            ub = 0.0f;
        }
        if (!GENERAL || (e_cur >= v.lo && e_cur < v.hi)) U[e_cur] = ua;
        if (has_b && (!GENERAL || (e_cur + plane >= v.lo && e_cur + plane < v.hi))) U[e_cur + plane] = ub;
        if (GENERAL && Gz) {
            Gz[e_cur] = 0.0f;
            if (has_b) Gz[e_cur + plane] = 0.0f;
        }
    }
}

template <class NET, bool GENERAL, bool STASH = false>
__global__ void __launch_bounds__(kThreads, 3) fwd_nodes_kernel(NodeView v, Tasks T) {
    pdl_trigger();
    // balanced contiguous runs: the (replica, strip, x plane) triples in that order are cut into gridDim.x * T.split equal
    // ranges dealt round-robin, so every CTA does the same number of plane iterations (+-1) and a thread keeps its
    // (y, z) for a long x march
    const int nx = v.x_end - v.x_begin;
    const int64_t per_rep = (int64_t)T.mblocks * nx, total = per_rep * v.nrep;
    const int64_t nranges = (int64_t)gridDim.x * T.split;
    for (int64_t rg = blockIdx.x; rg < nranges; rg += gridDim.x) {
        const int64_t hi = total * (rg + 1) / nranges;
        for (int64_t p = total * rg / nranges; p < hi;) {
            const int rep = (int)(p / per_rep);
            const int64_t q = p - rep * per_rep;
            const int mb = (int)(q / nx), xo = (int)(q - (int64_t)mb * nx);
            const int len = (int)min((int64_t)(nx - xo), hi - p);
            fwd_task<NET, GENERAL, STASH>(v, T.plane, v.xe + (size_t)rep * v.ex, v.ye + (size_t)rep * v.ey,
                                          v.ze + (size_t)rep * v.ez, v.side + rep * v.rep_nodes, v.U + rep * v.rep_nodes,
                                          mb * kThreads + (int)threadIdx.x, v.x_begin + xo, v.x_begin + xo + len,
                                          (GENERAL && v.Gz) ? v.Gz + rep * v.rep_nodes : nullptr);
            p += len;
        }
    }
}

// A2: E[c] = sum_q B[c][q] U[node_c + off_q] + B[c][27]   (discretization.py:464-513); clears gE
// (thread per crossed site; a warp-per-site variant with a shuffle reduction measured 14.9 us against 8.4 us)
__global__ void __launch_bounds__(128) extrap_kernel(nbm_shared_step_t s) {
    pdl_trigger();
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= s.n_crossed) return;
    int64_t e = s.c_node[c];
    // weights item-major [c][28] or, transposed for this access pattern, slot-major [28][n_crossed]
    const float* B = s.B_soa ? s.B_soa + c : s.B + c * 28;
    const int64_t bq = s.B_soa ? s.n_crossed : 1;
    int64_t sx = (int64_t)s.ey * s.ez, sy = s.ez;
    // all 55 loads are issued before the first use (register arrays): the kernel sits on the step's critical path on small
    // lattices, where its cost is memory round trips, not throughput
    float w[28], u[27];
#pragma unroll
    for (int q = 0; q < 28; ++q) w[q] = B[q * bq];
#pragma unroll
    for (int q = 0; q < 27; ++q) {
        int a = q % 3 - 1, b = (q / 3) % 3 - 1, cc = q / 9 - 1;
        u[q] = s.U[e + a * sx + b * sy + cc];
    }
    float acc = w[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc = fmaf(w[q], u[q], acc);
    s.E[c] = acc;
    s.gE[c] = 0.0f;
}

// B: residual rows, 7-point stencil on U (discretization.py:366-379 after division by diag).
// One thread per lattice node; the (y,z) plane is flattened so every warp reads 32 consecutive floats
// of each of the 7 weight arrays (SoA), x planes are blockIdx.y.
__global__ void __launch_bounds__(kThreads) residual_kernel(nbm_shared_step_t s) {
    int plane = s.ey * s.ez;
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int ix = blockIdx.y + 1;
    if (m >= plane) return;
    int iy = m / s.ez, iz = m - iy * s.ez;
    if (iy < 1 || iy >= s.ey - 1 || iz < 1 || iz >= s.ez - 1) return;
    int64_t sx = plane, sy = s.ez;
    int64_t ne = sx * s.ex;
    int64_t e = ix * sx + m;
    float u0 = s.U[e];
    float r = __ldg(s.w + e) * u0;
    r = fmaf(__ldg(s.w + 1 * ne + e), s.U[e - sx], r);
    r = fmaf(__ldg(s.w + 2 * ne + e), s.U[e + sx], r);
    r = fmaf(__ldg(s.w + 3 * ne + e), s.U[e - sy], r);
    r = fmaf(__ldg(s.w + 4 * ne + e), s.U[e + sy], r);
    r = fmaf(__ldg(s.w + 5 * ne + e), s.U[e - 1], r);
    r = fmaf(__ldg(s.w + 6 * ne + e), s.U[e + 1], r);
    if (s.nl) {
        r = fmaf(s.nl[e], nl_apply(s.nonlinear_m, s.nl_coef_m, u0), r);
        r = fmaf(s.nl[ne + e], nl_apply(s.nonlinear_p, s.nl_coef_p, u0), r);
    }
    s.R[e] = r - __ldg(s.rhs + e);
}

// B2: irregular rows: add the far-side (E) terms
__global__ void irregular_fwd_kernel(nbm_shared_step_t s) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= s.n_irr) return;
    float r = 0.0f;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        int32_t c = s.irr_c[q * 7 + k];
        if (c >= 0) r = fmaf(s.irr_wE[q * 7 + k], s.E[c], r);
    }
    uint8_t nlr = s.irr_nl[q];
    if (nlr) {
        float Ec = s.E[s.irr_c[q * 7]];
        r = fmaf(s.irr_nlw[q],
                 nlr == 1 ? nl_apply(s.nonlinear_m, s.nl_coef_m, Ec) : nl_apply(s.nonlinear_p, s.nl_coef_p, Ec), r);
    }
    const int64_t e = s.irr_point[q];
    if (s.faces) {
        // the whole row lives in the list: weights on the 7 sites, rhs; the U-part of the nonlinear term from nl
        const int64_t sx = (int64_t)s.ey * s.ez, sy = s.ez;
        const int64_t off[7] = {0, -sx, sx, -sy, sy, -1, 1};
        const float u0 = s.U[e];
#pragma unroll
        for (int k = 0; k < 7; ++k) r = fmaf(s.irr_wU[q * 7 + k], s.U[e + off[k]], r);
        if (s.nl) {
            const int64_t ne = sx * s.ex;
            r = fmaf(s.nl[e], nl_apply(s.nonlinear_m, s.nl_coef_m, u0), r);
            r = fmaf(s.nl[ne + e], nl_apply(s.nonlinear_p, s.nl_coef_p, u0), r);
        }
        s.R[e] = r - s.irr_rhs[q];
    } else {
        s.R[e] += r;
    }
}

// C1: G[x] = sum_k w_k[x - off_k] R[x - off_k]  (adjoint of the 7-point rows)
__global__ void __launch_bounds__(kThreads) adjoint_kernel(nbm_shared_step_t s) {
    int plane = s.ey * s.ez;
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int ix = blockIdx.y;
    if (m >= plane) return;
    int iy = m / s.ez, iz = m - iy * s.ez;
    int64_t sx = plane, sy = s.ez;
    int64_t ne = sx * s.ex;
    int64_t e = ix * sx + m;
    float r0 = s.R[e];
    float g = __ldg(s.w + e) * r0;
    // the point p = x + e_a has x as its "minus a" site (slot 1,3,5); p = x - e_a has x as slot 2,4,6
    if (ix + 1 < s.ex) g = fmaf(__ldg(s.w + 1 * ne + e + sx), s.R[e + sx], g);
    if (ix > 0) g = fmaf(__ldg(s.w + 2 * ne + e - sx), s.R[e - sx], g);
    if (iy + 1 < s.ey) g = fmaf(__ldg(s.w + 3 * ne + e + sy), s.R[e + sy], g);
    if (iy > 0) g = fmaf(__ldg(s.w + 4 * ne + e - sy), s.R[e - sy], g);
    if (iz + 1 < s.ez) g = fmaf(__ldg(s.w + 5 * ne + e + 1), s.R[e + 1], g);
    if (iz > 0) g = fmaf(__ldg(s.w + 6 * ne + e - 1), s.R[e - 1], g);
    if (s.nl) {
        float u0 = s.U[e];
        g = fmaf(s.nl[e] * nl_deriv(s.nonlinear_m, s.nl_coef_m, u0) +
                     s.nl[ne + e] * nl_deriv(s.nonlinear_p, s.nl_coef_p, u0), r0, g);
    }
    s.G[e] = g;
}


// ---------------------------------------------------------------------------------------------
// Vectorised forms of the two HBM-bound stencil kernels (used when the (y,z) plane is a multiple of 4
// and ez is even): one thread owns 4 consecutive cells of the flattened plane, the 7 weight arrays and
// rhs are read with 16-byte streaming loads (ld.global.cs: the 553 MB row table is touched once per
// kernel and must not evict U / R / G, which fit in the 126 MB L2), z neighbours come from the same
// registers, y neighbours from two 8-byte loads, x neighbours from aligned 16-byte loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld22(const float* p) {  // 8-byte aligned
    float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + 2);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

__global__ void __launch_bounds__(kThreads) residual4_kernel(nbm_shared_step_t s) {
    const int plane = s.ey * s.ez;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int ix = blockIdx.y + 1;
    if (m >= plane) return;
    const int64_t sx = plane, sy = s.ez;
    const int64_t ne = sx * s.ex;
    const int64_t e = ix * sx + m;
    // rows of the first / last y line and the z ends do not exist: their weights are zero in the table, and
    // the neighbour reads below stay inside the lattice because 1 <= ix <= ex-2 and the plane has a halo line.
    const bool first = (m == 0), last = (m + 4 >= plane);
    float4 u0 = ld4(s.U + e);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        float4 w = ldcs4(s.w + e);
        r = make_float4(w.x * u0.x, w.y * u0.y, w.z * u0.z, w.w * u0.w);
    }
    r = fma4(ldcs4(s.w + 1 * ne + e), ld4(s.U + e - sx), r);
    r = fma4(ldcs4(s.w + 2 * ne + e), ld4(s.U + e + sx), r);
    {
        float4 w3 = ldcs4(s.w + 3 * ne + e), w4 = ldcs4(s.w + 4 * ne + e);
        float4 um = (m >= sy) ? ld22(s.U + e - sy) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 up = (m + 4 + sy <= plane) ? ld22(s.U + e + sy) : make_float4(0.f, 0.f, 0.f, 0.f);
        r = fma4(w3, um, r);
        r = fma4(w4, up, r);
    }
    {
        float4 w5 = ldcs4(s.w + 5 * ne + e), w6 = ldcs4(s.w + 6 * ne + e);
        float ul = first ? 0.f : s.U[e - 1], ur = last ? 0.f : s.U[e + 4];
        r = fma4(w5, make_float4(ul, u0.x, u0.y, u0.z), r);
        r = fma4(w6, make_float4(u0.y, u0.z, u0.w, ur), r);
    }
    if (s.nl) {
        float4 a = ld4(s.nl + e), b = ld4(s.nl + ne + e);
        r.x += a.x * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.x) + b.x * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.x);
        r.y += a.y * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.y) + b.y * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.y);
        r.z += a.z * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.z) + b.z * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.z);
        r.w += a.w * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.w) + b.w * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.w);
    }
    float4 rh = ldcs4(s.rhs + e);
    *reinterpret_cast<float4*>(s.R + e) = make_float4(r.x - rh.x, r.y - rh.y, r.z - rh.z, r.w - rh.w);
}

__global__ void __launch_bounds__(kThreads) adjoint4_kernel(nbm_shared_step_t s) {
    const int plane = s.ey * s.ez;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int ix = blockIdx.y;
    if (m >= plane) return;
    const int64_t sx = plane, sy = s.ez;
    const int64_t ne = sx * s.ex;
    const int64_t e = ix * sx + m;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r0 = ld4(s.R + e);
    float4 w0 = ldcs4(s.w + e);
    float4 g = make_float4(w0.x * r0.x, w0.y * r0.y, w0.z * r0.z, w0.w * r0.w);
    // slot 1 (x-) of the row at x + e_x, slot 2 (x+) of the row at x - e_x, and so on
    if (ix + 1 < s.ex) g = fma4(ldcs4(s.w + 1 * ne + e + sx), ld4(s.R + e + sx), g);
    if (ix > 0) g = fma4(ldcs4(s.w + 2 * ne + e - sx), ld4(s.R + e - sx), g);
    if (m + 4 + sy <= plane) g = fma4(ld22(s.w + 3 * ne + e + sy), ld22(s.R + e + sy), g);
    if (m >= sy) g = fma4(ld22(s.w + 4 * ne + e - sy), ld22(s.R + e - sy), g);
    {
        // z neighbours: rows at x+1 (their slot 5) and x-1 (their slot 6); a row at the end of a y line has zero
        // weights toward the wrapped-around cell, so the flattened +-1 access is harmless
        const bool first = (m == 0), last = (m + 4 >= plane);
        float4 w5 = ldcs4(s.w + 5 * ne + e), w6 = ldcs4(s.w + 6 * ne + e);
        float w5r = last ? 0.f : s.w[5 * ne + e + 4], rr = last ? 0.f : s.R[e + 4];
        float w6l = first ? 0.f : s.w[6 * ne + e - 1], rl = first ? 0.f : s.R[e - 1];
        g = fma4(make_float4(w5.y, w5.z, w5.w, w5r), make_float4(r0.y, r0.z, r0.w, rr), g);
        g = fma4(make_float4(w6l, w6.x, w6.y, w6.z), make_float4(rl, r0.x, r0.y, r0.z), g);
    }
    if (s.nl) {
        float4 u0 = ld4(s.U + e), a = ld4(s.nl + e), b = ld4(s.nl + ne + e);
        g.x += (a.x * nl_deriv(s.nonlinear_m, s.nl_coef_m, u0.x) + b.x * nl_deriv(s.nonlinear_p, s.nl_coef_p, u0.x)) * r0.x;
        g.y += (a.y * nl_deriv(s.nonlinear_m, s.nl_coef_m, u0.y) + b.y * nl_deriv(s.nonlinear_p, s.nl_coef_p, u0.y)) * r0.y;
        g.z += (a.z * nl_deriv(s.nonlinear_m, s.nl_coef_m, u0.z) + b.z * nl_deriv(s.nonlinear_p, s.nl_coef_p, u0.z)) * r0.z;
        g.w += (a.w * nl_deriv(s.nonlinear_m, s.nl_coef_m, u0.w) + b.w * nl_deriv(s.nonlinear_p, s.nl_coef_p, u0.w)) * r0.w;
    }
    (void)z4;
    *reinterpret_cast<float4*>(s.G + e) = g;
}


// ---------------------------------------------------------------------------------------------
// "faces" table (nbm_assemble_t.faces): the finite-volume matrix of regular rows is symmetric, so one
// un-normalised coefficient per FACE (3 per node) + 1/diag replaces the 7 row weights: 28 B/node in the
// residual pass and 24 B/node in the adjoint pass instead of 40 / 36.  Irregular rows live entirely in
// the list, Dirichlet rows are dinv = -1.
//   regular row:  r = dinv * ( (sum_f c_f) u_p - sum_f c_f u_nb(f) ) - rhs          (discretization.py:366-386)
// ---------------------------------------------------------------------------------------------
struct Faces4 {
    float4 cxp, cxm, cyp, cym, czp, czm;
};
// T = dinv * R for regular rows (0 otherwise): d loss / d (un-normalised row)
__device__ __forceinline__ float4 tval4(float4 di, float4 r) {
    return make_float4(di.x > 0.f ? di.x * r.x : 0.f, di.y > 0.f ? di.y * r.y : 0.f, di.z > 0.f ? di.z * r.z : 0.f,
                       di.w > 0.f ? di.w * r.w : 0.f);
}
__device__ __forceinline__ float tval(float di, float r) { return di > 0.f ? di * r : 0.f; }
__device__ __forceinline__ Faces4 load_faces4(const nbm_shared_step_t& s, int64_t e, int m, int ix, int64_t sx,
                                              int64_t sy, int64_t ne, int plane) {
    Faces4 f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    f.cxp = ld4(s.cface + e);
    f.cxm = ix > 0 ? ld4(s.cface + e - sx) : z4;
    f.cyp = ld4(s.cface + ne + e);
    f.cym = m >= sy ? ld22(s.cface + ne + e - sy) : z4;
    f.czp = ld4(s.cface + 2 * ne + e);
    float l = (m == 0 && ix == 0) ? 0.f : s.cface[2 * ne + e - 1];
    f.czm = make_float4(l, f.czp.x, f.czp.y, f.czp.z);
    (void)plane;
    return f;
}
#define NBM_DIAG(F, c) ((((((F).cxm.c + (F).cxp.c) + (F).cym.c) + (F).cyp.c) + (F).czm.c) + (F).czp.c)

// d/du of the nonlinear part of a row: a N_m'(u) + b N_p'(u), in one fixed operation order (shared by the adjoint
// kernels so that their results agree bitwise)
__device__ __forceinline__ float nl_dfac(const nbm_shared_step_t& s, float a, float b, float u0) {
    return fmaf(a, nl_deriv(s.nonlinear_m, s.nl_coef_m, u0), __fmul_rn(b, nl_deriv(s.nonlinear_p, s.nl_coef_p, u0)));
}

// rows of the 4 consecutive cells m..m+3 of plane ix (1 <= ix <= ex-2)
__device__ __forceinline__ void residual_faces4_body(const nbm_shared_step_t& s, const int plane, const int m, const int ix) {
    const int64_t sx = plane, sy = s.ez;
    const int64_t ne = sx * s.ex;
    const int64_t e = ix * sx + m;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool first = (m == 0), last = (m + 4 >= plane);
    const float4 di = ldcs4(s.dinv + e);
    const float4 rh = ldcs4(s.rhs + e);
    const Faces4 F = load_faces4(s, e, m, ix, sx, sy, ne, plane);
    const float4 kv = s.kv ? ldcs4(s.kv + e) : z4;
    const float4 u0 = ld4(s.U + e);
    const float4 uxm = ld4(s.U + e - sx), uxp = ld4(s.U + e + sx);
    const float4 uym = (m >= sy) ? ld22(s.U + e - sy) : z4;
    const float4 uyp = (m + 4 + sy <= plane) ? ld22(s.U + e + sy) : z4;
    const float ul = first ? 0.f : s.U[e - 1], ur = last ? 0.f : s.U[e + 4];
    const float4 uzm = make_float4(ul, u0.x, u0.y, u0.z), uzp = make_float4(u0.y, u0.z, u0.w, ur);
    float4 r;
#define NBM_ROW(c)                                                                                                   \
    {                                                                                                                \
        float acc = (NBM_DIAG(F, c) + kv.c) * u0.c;                                                                         \
        acc = fmaf(-F.cxm.c, uxm.c, acc); acc = fmaf(-F.cxp.c, uxp.c, acc);                                          \
        acc = fmaf(-F.cym.c, uym.c, acc); acc = fmaf(-F.cyp.c, uyp.c, acc);                                          \
        acc = fmaf(-F.czm.c, uzm.c, acc); acc = fmaf(-F.czp.c, uzp.c, acc);                                          \
        r.c = di.c > 0.f ? fmaf(di.c, acc, -rh.c) : (di.c < 0.f ? u0.c - rh.c : 0.f);                                \
    }
    NBM_ROW(x) NBM_ROW(y) NBM_ROW(z) NBM_ROW(w)
#undef NBM_ROW
    if (s.nl) {
        // (rows without a dense row - irregular rows live in the list - stay exactly 0 here)
        float4 a = ld4(s.nl + e), b = ld4(s.nl + ne + e);
        if (di.x != 0.f) r.x += a.x * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.x) + b.x * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.x);
        if (di.y != 0.f) r.y += a.y * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.y) + b.y * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.y);
        if (di.z != 0.f) r.z += a.z * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.z) + b.z * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.z);
        if (di.w != 0.f) r.w += a.w * nl_apply(s.nonlinear_m, s.nl_coef_m, u0.w) + b.w * nl_apply(s.nonlinear_p, s.nl_coef_p, u0.w);
    }
    *reinterpret_cast<float4*>(s.R + e) = r;
    if (s.S) *reinterpret_cast<float4*>(s.S + e) = tval4(di, r);
}

__global__ void __launch_bounds__(kThreads) residual_faces4_kernel(nbm_shared_step_t s) {
    const int plane = s.ey * s.ez;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (m >= plane) return;
    residual_faces4_body(s, plane, m, blockIdx.y + 1);
}

// G of the 4 consecutive cells m..m+3 of plane ix (0 <= ix <= ex-1)
__device__ __forceinline__ void adjoint_faces4_body(const nbm_shared_step_t& s, const int plane, const int m, const int ix) {
    const int64_t sx = plane, sy = s.ez;
    const int64_t ne = sx * s.ex;
    const int64_t e = ix * sx + m;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool first = (m == 0 && ix == 0), last = (m + 4 >= plane && ix + 1 >= s.ex);
    const float4 di = ldcs4(s.dinv + e);
    const float4 r0 = ld4(s.R + e);
    const Faces4 F = load_faces4(s, e, m, ix, sx, sy, ne, plane);
    const float4 t0 = tval4(di, r0);
    const float4 txm = ix > 0 ? tval4(ld4(s.dinv + e - sx), ld4(s.R + e - sx)) : z4;
    const float4 txp = ix + 1 < s.ex ? tval4(ld4(s.dinv + e + sx), ld4(s.R + e + sx)) : z4;
    const float4 tym = m >= sy ? tval4(ld22(s.dinv + e - sy), ld22(s.R + e - sy)) : z4;
    const float4 typ = m + 4 + sy <= plane ? tval4(ld22(s.dinv + e + sy), ld22(s.R + e + sy)) : z4;
    const float tl = first ? 0.f : tval(s.dinv[e - 1], s.R[e - 1]);
    const float tr = last ? 0.f : tval(s.dinv[e + 4], s.R[e + 4]);
    const float4 tzm = make_float4(tl, t0.x, t0.y, t0.z), tzp = make_float4(t0.y, t0.z, t0.w, tr);
    const float4 kv = s.kv ? ldcs4(s.kv + e) : z4;
    float4 g;
#define NBM_ADJ(c)                                                                                                   \
    {                                                                                                                \
        float acc = di.c > 0.f ? (NBM_DIAG(F, c) + kv.c) * t0.c : (di.c < 0.f ? r0.c : 0.f);                                 \
        acc = fmaf(-F.cxm.c, txm.c, acc); acc = fmaf(-F.cxp.c, txp.c, acc);                                          \
        acc = fmaf(-F.cym.c, tym.c, acc); acc = fmaf(-F.cyp.c, typ.c, acc);                                          \
        acc = fmaf(-F.czm.c, tzm.c, acc); acc = fmaf(-F.czp.c, tzp.c, acc);                                          \
        g.c = acc;                                                                                                   \
    }
    NBM_ADJ(x) NBM_ADJ(y) NBM_ADJ(z) NBM_ADJ(w)
#undef NBM_ADJ
    if (s.nl) {
        float4 u0 = ld4(s.U + e), a = ld4(s.nl + e), b = ld4(s.nl + ne + e);
        g.x = fmaf(nl_dfac(s, a.x, b.x, u0.x), r0.x, g.x);
        g.y = fmaf(nl_dfac(s, a.y, b.y, u0.y), r0.y, g.y);
        g.z = fmaf(nl_dfac(s, a.z, b.z, u0.z), r0.z, g.z);
        g.w = fmaf(nl_dfac(s, a.w, b.w, u0.w), r0.w, g.w);
    }
    *reinterpret_cast<float4*>(s.G + e) = g;
}

__global__ void __launch_bounds__(kThreads) adjoint_faces4_kernel(nbm_shared_step_t s) {
    const int plane = s.ey * s.ez;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (m >= plane) return;
    adjoint_faces4_body(s, plane, m, blockIdx.y);
}

// C0 (deterministic form): the adjoint of the lists GATHERED through the transposed incidence (CSR built once per
// level on the host): no atomics, fixed summation order.
//   gE[c] = sum over (row q, slot k) with irr_c[q][k] == c of  wE[q][k] R[q]  (+ the nonlinear term of slot 0)
__global__ void gather_gE_kernel(nbm_shared_step_t s) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= s.n_crossed) return;
    float acc = 0.0f;
    const int b = s.ge_ptr[c], e = s.ge_ptr[c + 1];
    for (int i = b; i < e; ++i) {
        const int ent = s.ge_ent[i];
        const int q = ent >> 3, k = ent & 7;
        const float r = s.R[s.irr_point[q]];
        acc = fmaf(s.irr_wE[(int64_t)q * 7 + k], r, acc);
        if (k == 0) {
            const uint8_t nlr = s.irr_nl[q];
            if (nlr) {
                const float Ec = s.E[c];
                const float d = nlr == 1 ? nl_deriv(s.nonlinear_m, s.nl_coef_m, Ec) : nl_deriv(s.nonlinear_p, s.nl_coef_p, Ec);
                acc = fmaf(s.irr_nlw[q] * d, r, acc);
            }
        }
    }
    s.gE[c] = acc;
}
//   G[n] += sum of irr_wU[q][k] R[q] over the irregular rows that use node n (faces table)
//         + sum of B[c][v] gE[c] over the crossed sites whose 27-cube holds node n
__global__ void gather_G_kernel(nbm_shared_step_t s) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n_list) return;
    float acc = 0.0f;
    const int b = s.g_ptr[i], e = s.g_ptr[i + 1];
    for (int j = b; j < e; ++j) {
        const int ent = s.g_ent[j];
        if (ent >= 0) {
            const int q = ent >> 3, k = ent & 7;
            acc = fmaf(s.irr_wU[(int64_t)q * 7 + k], s.R[s.irr_point[q]], acc);
        } else {
            const int m = -(ent + 1);
            const int c = m >> 5, v = m & 31;
            acc = fmaf(s.B[(int64_t)c * 28 + v], s.gE[c], acc);
        }
    }
    s.G[s.list_nodes[i]] += acc;
}

// C0: adjoint of the irregular rows: gE[c] += wE * R[p].  `nl_center`: the dense adjoint ran BEFORE this row's
// residual existed (fused dense stage), so the U-part of the row's nonlinear term is added here instead.
__global__ void irregular_bwd_kernel(nbm_shared_step_t s, bool nl_center) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= s.n_irr) return;
    const int64_t e = s.irr_point[q];
    float r = s.R[e];
    if (nl_center && s.nl) {
        const int64_t ne = (int64_t)s.ex * s.ey * s.ez;
        const float u0 = s.U[e];
        atomicAdd(s.G + e, (s.nl[e] * nl_deriv(s.nonlinear_m, s.nl_coef_m, u0) +
                            s.nl[ne + e] * nl_deriv(s.nonlinear_p, s.nl_coef_p, u0)) * r);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        int32_t c = s.irr_c[q * 7 + k];
        if (c >= 0) atomicAdd(s.gE + c, s.irr_wE[q * 7 + k] * r);
    }
    if (s.faces) {
        const int64_t sx = (int64_t)s.ey * s.ez, sy = s.ez;
        const int64_t off[7] = {0, -sx, sx, -sy, sy, -1, 1};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            float w = s.irr_wU[q * 7 + k];
            if (w != 0.0f) atomicAdd(s.G + e + off[k], w * r);
        }
        // (the U-part of the nonlinear term is in the dense nl table: the dense adjoint pass already added it)
    }
    uint8_t nlr = s.irr_nl[q];
    if (nlr) {
        int32_t c0 = s.irr_c[q * 7];
        float Ec = s.E[c0];
        float d = nlr == 1 ? nl_deriv(s.nonlinear_m, s.nl_coef_m, Ec) : nl_deriv(s.nonlinear_p, s.nl_coef_p, Ec);
        atomicAdd(s.gE + c0, s.irr_nlw[q] * d * r);
    }
}

// C0b: adjoint of the extrapolation: G[node_c + off_q] += B[c][q] gE[c]
__global__ void extrap_bwd_kernel(nbm_shared_step_t s, float* __restrict__ Gt) {
    pdl_wait();
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t c = t / 27;
    int q = (int)(t - c * 27);
    if (c >= s.n_crossed) return;
    float g = s.gE[c];
    if (g == 0.0f) return;
    int64_t sx = (int64_t)s.ey * s.ez, sy = s.ez;
    int a = q % 3 - 1, b = (q / 3) % 3 - 1, cc = q / 9 - 1;
    atomicAdd(Gt + s.c_node[c] + a * sx + b * sy + cc, s.B[c * 28 + q] * g);   // (27 consecutive threads: item-major is coalesced)
}

// ---- the list chain beside the dense stencil (nbm_shared_step_t.G2 / Rq; faces table) -----------------------------
// An irregular row forward AND backward in one thread: the row needs only U and E, its residual stays in the compact
// array Rq (the dense kernel owns R while this runs), its adjoint goes into gE and into the side buffer G2.
__global__ void irregular_fb_kernel(nbm_shared_step_t s) {
    pdl_trigger();
    pdl_wait();
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= s.n_irr) return;
    const int64_t e = s.irr_point[q];
    const int64_t sx = (int64_t)s.ey * s.ez, sy = s.ez;
    const int64_t off[7] = {0, -sx, sx, -sy, sy, -1, 1};
    int32_t c[7];
    float wE[7], wU[7], u[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        if (s.irr_c_soa) {     // slot-major copies: consecutive rows of a warp read consecutive words
            c[k] = s.irr_c_soa[k * s.n_irr + q];
            wE[k] = s.irr_wE_soa[k * s.n_irr + q];
            wU[k] = s.irr_wU_soa[k * s.n_irr + q];
        } else {
            c[k] = s.irr_c[q * 7 + k];
            wE[k] = s.irr_wE[q * 7 + k];
            wU[k] = s.irr_wU[q * 7 + k];
        }
        u[k] = s.U[e + off[k]];
    }
    float Ek[7];          // (the far-side values are fetched together, before the accumulation chain needs the first)
#pragma unroll
    for (int k = 0; k < 7; ++k) Ek[k] = c[k] >= 0 ? s.E[c[k]] : 0.0f;
    const float rhs_q = s.irr_rhs[q];
    // same operation order as irregular_fwd_kernel
    float r = 0.0f;
#pragma unroll
    for (int k = 0; k < 7; ++k)
        if (c[k] >= 0) r = fmaf(wE[k], Ek[k], r);
    const uint8_t nlr = s.irr_nl[q];
    float Ec = 0.0f, nlw = 0.0f;
    if (nlr) {
        Ec = Ek[0];
        nlw = s.irr_nlw[q];
        r = fmaf(nlw, nlr == 1 ? nl_apply(s.nonlinear_m, s.nl_coef_m, Ec) : nl_apply(s.nonlinear_p, s.nl_coef_p, Ec), r);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) r = fmaf(wU[k], u[k], r);
    float nla = 0.0f, nlb = 0.0f;
    if (s.nl) {
        const int64_t ne = sx * s.ex;
        nla = s.nl[e];
        nlb = s.nl[ne + e];
        r = fmaf(nla, nl_apply(s.nonlinear_m, s.nl_coef_m, u[0]), r);
        r = fmaf(nlb, nl_apply(s.nonlinear_p, s.nl_coef_p, u[0]), r);
    }
    r -= rhs_q;
    s.Rq[q] = r;
    // adjoint (irregular_bwd_kernel with nl_center = true)
    if (s.nl)
        atomicAdd(s.G2 + e, (nla * nl_deriv(s.nonlinear_m, s.nl_coef_m, u[0]) + nlb * nl_deriv(s.nonlinear_p, s.nl_coef_p, u[0])) * r);
#pragma unroll
    for (int k = 0; k < 7; ++k)
        if (c[k] >= 0) atomicAdd(s.gE + c[k], wE[k] * r);
#pragma unroll
    for (int k = 0; k < 7; ++k)
        if (wU[k] != 0.0f) atomicAdd(s.G2 + e + off[k], wU[k] * r);
    if (nlr) {
        const float d = nlr == 1 ? nl_deriv(s.nonlinear_m, s.nl_coef_m, Ec) : nl_deriv(s.nonlinear_p, s.nl_coef_p, Ec);
        atomicAdd(s.gE + c[0], nlw * d * r);
    }
}

// join of the list chain: G += G2 on the nodes the lists can reach (G2 re-zeroed), R <- Rq on the irregular rows
__global__ void merge_lists_kernel(nbm_shared_step_t s) {
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < s.n_list) {
        const int64_t n = s.list_nodes[i];
        const float g2 = s.G2[n];
        if (g2 != 0.0f) {
            s.G[n] += g2;
            s.G2[n] = 0.0f;
        }
    }
    if (i < s.n_irr) s.R[s.irr_point[i]] = s.Rq[i];
}

// block-level reduction of per-thread accumulators into partials[blockIdx.x][0..NP] (loss last)
template <class NET>
__device__ __forceinline__ void block_reduce_store(typename NET::Acc& acc, float loss, float* __restrict__ partials,
                                                   int stride = NET::NP + 1) {
    constexpr int NP = NET::NP;
    __shared__ float sm[kThreads / 32][NP + 1];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i <= NP; ++i) {
        float v = i < NP ? acc.get(i) : loss;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP + 1; i += kThreads) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += sm[w][i];
        partials[(size_t)blockIdx.x * stride + (i < NP ? i : stride - 1)] = v;
    }
}

// C2: d loss/d theta = sum_nodes (G/n) d u(node)/d theta, loss = sum_rows 0.5 R^2 / n
// (value_and_grad(self.loss), trainer.py:786; mean of optax.l2_loss, :899-901)
// GENERAL = false: the shared path (one replica, every node of the lattice, loss from R);
// GENERAL = true : the 7 displaced lattices of the general path (node range filter, no R)
//
// 12 warps per SM (168 registers): the p-head accumulators are pair-split (MlpP::AccS), the y/z rows and
// the bias of the first layer live in shared memory and are touched once per task.
#ifndef NBM_GRAD_THREADS
#define NBM_GRAD_THREADS 384
#endif
constexpr int kGradThreads = NBM_GRAD_THREADS;

constexpr int kStashStages = 3;   // cp.async ring of the activation stash: nodes ix, ix+1, ix+2
template <class NET>
constexpr int grad_smem_bytes(bool stash = false) {
    return (3 * NET::HPW * kGradThreads + (kGradThreads / 32) * (NET::NP + 1)) * (int)sizeof(float) +
           (stash ? kStashStages * NET::P::HPQ * kGradThreads * 16 + 16 : 0);
}

template <class NET, bool GENERAL, bool STASH = false>
__global__ void __launch_bounds__(kGradThreads, 1) node_grad_kernel(NodeView v, Tasks T) {
    using P = typename NET::P;
    using M = typename NET::M;
    constexpr int H = NET::HPW, HP2 = H / 2, NP = NET::NP;
    extern __shared__ __align__(16) float dsm[];
    float* hs = dsm + threadIdx.x;                      // [3H] hoisted sums of this thread, stride kGradThreads
    float* red = dsm + 3 * H * kGradThreads;            // [warps][NP + 1]
    // activation stash ring: [stage][thread][HPQ] 16-byte chunks filled by this thread's own cp.async copies, two nodes
    // ahead of the compute (no register is held by a load in flight, DRAM latency is off the critical path)
    constexpr int HPQ = P::HPQ;
    const uint32_t hsm_a = STASH ? (((uint32_t)__cvta_generic_to_shared(red + (kGradThreads / 32) * (NP + 1)) + 15u) & ~15u) +
                                       16u * HPQ * threadIdx.x : 0u;
    constexpr uint32_t kStageB = HPQ * kGradThreads * 16u;
#pragma unroll
    for (int i = 0; i < 3 * H; ++i) hs[i * kGradThreads] = 0.0f;
    typename P::AccS acc;
    acc.zero();
    float accm[M::NP];
#pragma unroll
    for (int i = 0; i < M::NP; ++i) accm[i] = 0.0f;
    float loss = 0.0f;
    const bool par = (threadIdx.x & 1) != 0;
    const float* R = v.R;  // only the shared path (one replica) accumulates the loss here
    pdl_trigger();
    pdl_wait();            // (programmatic launch: G, R of the previous kernel are complete from here on)
    // balanced contiguous runs of (replica, strip, x plane) triples (see fwd_nodes_kernel)
    const int nx = v.x_end - v.x_begin;
    const int64_t per_rep = (int64_t)T.mblocks * nx, total = per_rep * v.nrep;
    const int64_t nranges = (int64_t)gridDim.x * T.split;
    int64_t rg = blockIdx.x, pr = total * rg / nranges, run_hi = total * (rg + 1) / nranges;
    for (;;) {
        if (pr >= run_hi) {        // next range of this CTA
            rg += gridDim.x;
            if (rg >= nranges) break;
            pr = total * rg / nranges;
            run_hi = total * (rg + 1) / nranges;
            continue;
        }
        const int rep = (int)(pr / per_rep);
        const int64_t qr = pr - rep * per_rep;
        const int mb = (int)(qr / nx), xo = (int)(qr - (int64_t)mb * nx);
        const int run_len = (int)min((int64_t)(nx - xo), run_hi - pr);
        pr += run_len;
        const float* xe = v.xe + (size_t)rep * v.ex;
        const float* ye = v.ye + (size_t)rep * v.ey;
        const float* ze = v.ze + (size_t)rep * v.ez;
        const uint8_t* side = v.side + rep * v.rep_nodes;
        const float* G = v.G + rep * v.rep_nodes;
        const int m_raw = mb * kGradThreads + threadIdx.x;
        const bool valid = m_raw < T.plane;             // lanes past the plane stay in the warp (shuffles) with g = 0
        const int m = valid ? m_raw : T.plane - 1;
        const int iy = m / v.ez, iz = m - iy * v.ez;
        const float y = __ldg(ye + iy), z = __ldg(ze + iz);
        u64 yz[HP2];
        P::template first_layer_yz<0>(y, z, yz);
        const int x0 = v.x_begin + xo, x1 = x0 + run_len;
        // software pipeline: the loads of plane ix+1 are in flight while plane ix is computed; running pointers, one
        // add per array and plane
        int64_t e = (int64_t)x0 * T.plane + m;
        const float* gp = G + e;
        const float* rp = (GENERAL || !R) ? nullptr : R + e;
        const uint8_t* sp = side + e;
        const float* xp = xe + x0;
        bool in_n = valid && (!GENERAL || (e >= v.lo && e < v.hi));
        float g_n = in_n ? __ldg(gp) : 0.0f, r_n = (rp && valid) ? __ldg(rp) : 0.0f, x_n = __ldg(xp);
        uint8_t sd_n = __ldg(sp);
        const float4* hnext = STASH ? v.Hst + e : nullptr;   // next node to prefetch from the stash
        uint32_t st_w = 0, st_r = 0;                      // ring stage written next / read next
        auto stash_prefetch = [&]() {
            int64_t hs16 = v.rep_nodes;
            asm volatile("" : "+l"(hs16));    // opaque: one running pointer + stride instead of HPQ hoisted pointers
#pragma unroll
            for (int q = 0; q < HPQ; ++q)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(hsm_a + st_w * kStageB + 16u * q),
                             "l"(hnext + q * hs16) : "memory");
            hnext += T.plane;
            st_w = st_w + 1 == kStashStages ? 0u : st_w + 1;
        };
        if (STASH) {
            stash_prefetch();
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (x0 + 1 < x1) stash_prefetch();
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int ix = x0; ix < x1; ++ix) {
            const float g = g_n * v.inv_n, r = r_n, x = x_n;
            const bool plus = (sd_n & 1) != 0;
            if (ix + 1 < x1) {
                gp += T.plane;
                sp += T.plane;
                ++xp;
                if (GENERAL) {
                    e += T.plane;
                    in_n = valid && e >= v.lo && e < v.hi;
                }
                g_n = in_n ? __ldg(gp) : 0.0f;
                if (!GENERAL && rp) {
                    rp += T.plane;
                    r_n = valid ? __ldg(rp) : 0.0f;
                }
                sd_n = __ldg(sp);
                x_n = __ldg(xp);
            }
            loss = fmaf(0.5f * r, r, loss);
            const bool do_p = plus && g != 0.0f;
            u64 aLast[HP2];
            if (STASH) {
                if (ix + 2 < x1) stash_prefetch();
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 2;" ::: "memory");     // this node's copies have landed
#pragma unroll
                for (int q = 0; q < HPQ; ++q) {
                    float4 c;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w)
                                 : "r"(hsm_a + st_r * kStageB + 16u * q));
                    aLast[2 * q] = pk(c.x, c.y);
                    if (2 * q + 1 < HP2) aLast[2 * q + 1] = pk(c.z, c.w);
                }
                st_r = st_r + 1 == kStashStages ? 0u : st_r + 1;
            }
            if (__any_sync(0xffffffffu, do_p)) P::template grad_split<STASH>(x, yz, do_p ? g : 0.0f, acc, par, aLast);
            if (!plus && g != 0.0f) {
                float a[NET::LMD][NET::HMW];
                M::template forward<P::NP>(x, y, z, a);
                M::template backward<P::NP, M::NP, P::NP>(x, y, z, a, g, accm);
            }
        }
        // fold the task's sum(delta1) into the bias and the y, z rows of the first layer
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float t = (j & 1) ? hi32(acc.t[j / 2]) : lo32(acc.t[j / 2]);
            hs[j * kGradThreads] += t;
            hs[(H + j) * kGradThreads] = fmaf(y, t, hs[(H + j) * kGradThreads]);
            hs[(2 * H + j) * kGradThreads] = fmaf(z, t, hs[(2 * H + j) * kGradThreads]);
        }
#pragma unroll
        for (int j = 0; j < HP2; ++j) acc.t[j] = 0ull;
    }
    loss *= v.inv_n;
    // block reduction into partials[row][0..NP] (loss last)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i <= NP; ++i) {
        float val;
        if (i < P::NP) val = P::split_get(acc, i, par, hs, kGradThreads);
        else if (i < NP) val = accm[i < NP ? i - P::NP : 0];
        else val = loss;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) red[warp * (NP + 1) + i] = val;
    }
    __syncthreads();
    const int stride = v.row_stride ? v.row_stride : NP + 1, loss_col = v.row_stride ? v.loss_col : NP;
    float* partials = v.partials + (size_t)v.row0 * stride;
    for (int i = threadIdx.x; i < NP + 1; i += kGradThreads) {
        float val = 0.0f;
#pragma unroll
        for (int w = 0; w < kGradThreads / 32; ++w) val += red[w * (NP + 1) + i];
        partials[(size_t)blockIdx.x * stride + (i < NP ? i : loss_col)] = val;
    }
}

template <class NET, bool GENERAL, bool STASH = false>
static cudaError_t launch_node_grad(dim3 grid, const NodeView& v, const Tasks& T, cudaStream_t st) {
    static unsigned long long configured = 0ull;   // per instantiation AND per device (one process may drive several)
    constexpr int bytes = grad_smem_bytes<NET>(STASH);
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(node_grad_kernel<NET, GENERAL, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        configured |= 1ull << (dev & 63);
    }
    return launch_pdl(node_grad_kernel<NET, GENERAL, STASH>, grid, dim3(kGradThreads), (size_t)bytes, st, v, T);
}

// ---------------------------------------------------------------------------------------------
// Fused adjoint + gradient (faces mode, nbm_shared_step_t.S != NULL).  The dense adjoint stencil
//   G[e] = D_e S[e] - sum_f c_f S[nb_f(e)],   S = dinv * R,
// is evaluated inside the gradient kernel.  A CTA walks (strip of 384 cells) x (x planes); for every plane
// one elected thread issues bulk async copies (cp.async.bulk, completion on an mbarrier) of the strip's
// windows of S (with its y/z halo), the three face-coefficient arrays, dinv, R, kv and side into a
// shared-memory ring, kRing - 2 planes ahead of the compute, so no register is held by a load in flight.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}

constexpr unsigned kRing = 8;    // stages of the shared-memory ring
constexpr unsigned kAhead = 4;   // loads in flight ahead of a warp's position; the other kRing - kAhead - 1 stages let
                                 // the warps of a CTA drift apart without blocking the issuer

struct FusedView {
    const float *xe, *ye, *ze;
    int ex, ey, ez;
    int64_t ne;
    const uint8_t* side;
    const float *S, *cface, *dinv, *kv, *R;
    float* G;                   // list contributions only; consumed entries are re-zeroed
    float inv_n;
    float* partials;
};

// shared-memory ring stage (offsets in floats, all multiples of 4 -> 16-byte aligned)
struct StageLayout {
    int ezu, oS, oCY, oCZ, oCX, oDI, oR, oKV, oSD, floats;
};
__host__ __device__ inline StageLayout stage_layout(int ez) {
    StageLayout L;
    L.ezu = (ez + 3) & ~3;
    L.oS = 0;
    L.oCY = L.oS + kGradThreads + 2 * L.ezu;
    L.oCZ = L.oCY + kGradThreads + L.ezu;
    L.oCX = L.oCZ + kGradThreads + 4;
    L.oDI = L.oCX + kGradThreads;
    L.oR = L.oDI + kGradThreads;
    L.oKV = L.oR + kGradThreads;
    L.oSD = L.oKV + kGradThreads;
    L.floats = L.oSD + (kGradThreads + 32) / 4;
    return L;
}
template <class NET>
static int fused_smem_bytes(int ez) {
    return grad_smem_bytes<NET>() + 16 + kRing * stage_layout(ez).floats * (int)sizeof(float);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}

// what lane a < 8 of the issuing warp copies for every plane
struct CopyDesc {
    const char* base;   // array base
    int64_t limit;      // elements readable
    int lo_off, hi_off; // window [e0 + lo_off, e0 + hi_off) in elements
    int dst_bytes;      // offset of the window inside a stage
    int esz;            // element size; 1: the window is widened to 16-element boundaries
};

// raw shared-window addresses (the generic -> shared conversion is done once per kernel)
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// Every task stages the same number of planes, x0-1 .. x0+xchunk (planes outside the lattice or past the chunk are
// "empty loads": the barrier completes with zero bytes), so load n <-> (task, plane) is closed-form:
//   n = k * (xchunk + 2) + j,  task = blockIdx.x + k * gridDim.x,  plane = x0 - 1 + j.
// Load n is issued by warp n % kWarps when that warp reaches position n - kAhead (lanes 0..7 copy one window each).
template <class NET>
__global__ void __launch_bounds__(kGradThreads, 1) node_grad_fused_kernel(FusedView v, Tasks T) {
    using P = typename NET::P;
    using M = typename NET::M;
    constexpr int H = NET::HPW, HP2 = H / 2, NP = NET::NP;
    constexpr int kRed = ((kGradThreads / 32) * (NP + 1) + 3) & ~3;
    constexpr unsigned kWarps = kGradThreads / 32;
    extern __shared__ __align__(16) float dsm[];
    __shared__ __align__(8) uint64_t bars[2 * kRing];   // [0, kRing): full, [kRing, 2 kRing): empty
    __shared__ CopyDesc desc[8];
    float* hs = dsm + threadIdx.x;
    float* red = dsm + 3 * H * kGradThreads;
    float* ring = red + kRed;
    const StageLayout SL = stage_layout(v.ez);
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31, warp = tid >> 5;
    const int plane = T.plane;
    const int64_t ne = v.ne;
    const uint32_t full_a = smem_u32(bars), empty_a = full_a + 8u * kRing;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < (int)kRing; ++s) {
            mbar_init(&bars[s], 8);               // the 8 issuing lanes arrive (each with its own byte count)
            mbar_init(&bars[kRing + s], kWarps);  // every warp releases a stage after its last read
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int64_t ne16 = (ne + 15) & ~(int64_t)15;
        const int G = kGradThreads;
        desc[0] = {(const char*)v.S, ne, -SL.ezu, G + SL.ezu, SL.oS * 4, 4};
        desc[1] = {(const char*)(v.cface + ne), ne, -SL.ezu, G, SL.oCY * 4, 4};
        desc[2] = {(const char*)(v.cface + 2 * ne), ne, -4, G, SL.oCZ * 4, 4};
        desc[3] = {(const char*)v.cface, ne, 0, G, SL.oCX * 4, 4};
        desc[4] = {(const char*)v.dinv, ne, 0, G, SL.oDI * 4, 4};
        desc[5] = {(const char*)v.R, ne, 0, G, SL.oR * 4, 4};
        desc[6] = {(const char*)v.kv, v.kv ? ne : 0, 0, G, SL.oKV * 4, 4};
        desc[7] = {(const char*)v.side, ne16, 0, G, SL.oSD * 4, 1};
    }
#pragma unroll
    for (int i = 0; i < 3 * H; ++i) hs[i * kGradThreads] = 0.0f;
    typename P::AccS acc;
    acc.zero();
    float accm[M::NP];
#pragma unroll
    for (int i = 0; i < M::NP; ++i) accm[i] = 0.0f;
    float loss = 0.0f;
    const bool par = (tid & 1) != 0;

    const unsigned per_task = (unsigned)T.xchunk + 2u;
    const unsigned my_tasks = blockIdx.x < (unsigned)T.total ? ((unsigned)T.total - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
    const unsigned n_total = my_tasks * per_task;
    // issue load n (whole warp, converged)
    auto issue = [&](unsigned n) {
        const unsigned st = n % kRing;
        if (n >= kRing) mbar_wait_a(empty_a + 8u * st, ((n / kRing) + 1u) & 1u);   // every warp released its previous use
        if (lane < 8) {
            const unsigned k = n / per_task, j = n - k * per_task;
            const int task = blockIdx.x + k * gridDim.x;
            const int mb = task % T.mblocks, xc = task / T.mblocks;
            const int x0 = xc * T.xchunk, x1 = min(v.ex, x0 + T.xchunk);
            const int q = x0 - 1 + (int)j;
            uint32_t bytes = 0u;
            const CopyDesc d = desc[lane];
            int64_t lo = 0, l = 0;
            if (q >= 0 && q <= x1 && q < v.ex) {
                const int64_t e0 = (int64_t)q * plane + mb * kGradThreads;
                lo = e0 + d.lo_off;
                int64_t hi = e0 + d.hi_off;
                if (d.esz == 1) {
                    lo &= ~(int64_t)15;
                    hi = (hi + 15) & ~(int64_t)15;
                }
                l = max(lo, (int64_t)0);
                const int64_t h = min(hi, d.limit);
                bytes = h > l ? (uint32_t)(h - l) * (uint32_t)d.esz : 0u;
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_a + 8u * st), "r"(bytes) : "memory");
            if (bytes)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(ring + st * SL.floats) + (uint32_t)d.dst_bytes + (uint32_t)((l - lo) * d.esz)),
                             "l"(d.base + l * d.esz), "r"(bytes), "r"(full_a + 8u * st)
                             : "memory");
        }
        __syncwarp();
    };
    __syncthreads();   // barriers and descriptors initialised
    for (unsigned n = 0; n < kAhead && n < n_total; ++n)
        if (n % kWarps == warp) issue(n);

    unsigned n_cur = 0;
    for (int task = blockIdx.x; task < T.total; task += gridDim.x) {
        const int mb = task % T.mblocks, xc = task / T.mblocks;
        const int m0 = mb * kGradThreads, x0 = xc * T.xchunk, x1 = min(v.ex, x0 + T.xchunk);
        const int m_raw = m0 + tid;
        const bool valid = m_raw < plane;
        const int m = valid ? m_raw : plane - 1;
        const int t = m - m0;
        const int iy = m / v.ez, iz = m - iy * v.ez;
        const float y = __ldg(v.ye + iy), z = __ldg(v.ze + iz);
        u64 yz[HP2];
        P::template first_layer_yz<0>(y, z, yz);
        // neighbours that fall outside the arrays (first rows of plane 0, last rows of plane ex-1) read as zero
        const bool lo_y = m < v.ez, hi_y = m >= plane - v.ez, lo_z = m == 0, hi_z = m == plane - 1;
        float S_prev = 0.0f, cxm = 0.0f;
        // positions j = 0 .. xchunk of the task (j = xchunk + 1, the plane after the chunk, is only read as "next")
        for (int j = 0; j <= T.xchunk; ++j, ++n_cur) {
            {
                const unsigned ni = n_cur + kAhead;
                if (ni < n_total && ni % kWarps == warp) issue(ni);
            }
            const int q = x0 - 1 + j;
            const unsigned st = n_cur % kRing;
            const float* cur = ring + st * SL.floats + t;
            const bool live = q >= 0 && q < x1;     // the plane exists and belongs to the chunk (or precedes it)
            if (live) mbar_wait_a(full_a + 8u * st, (n_cur / kRing) & 1u);
            if (j == 0 || !live) {
                if (live) {      // the plane before the chunk only feeds the x- neighbour
                    S_prev = cur[SL.oS + SL.ezu];
                    cxm = cur[SL.oCX];
                }
                __syncwarp();
                if (lane == 0) {
                    if (!live) mbar_wait_a(full_a + 8u * st, (n_cur / kRing) & 1u);   // (completes with zero bytes)
                    mbar_arrive_a(empty_a + 8u * st);
                    if (j == T.xchunk) {   // short last chunk: the task's final stage has no reader either
                        const unsigned s1 = (n_cur + 1u) % kRing;
                        mbar_wait_a(full_a + 8u * s1, ((n_cur + 1u) / kRing) & 1u);
                        mbar_arrive_a(empty_a + 8u * s1);
                    }
                }
                continue;
            }
            const float S0 = cur[SL.oS + SL.ezu];
            const float cxp = cur[SL.oCX];
            const unsigned nn = n_cur + 1u, sn = nn % kRing;
            float S_next = 0.0f;
            if (q + 1 < v.ex && q + 1 <= x1) {
                mbar_wait_a(full_a + 8u * sn, (nn / kRing) & 1u);
                S_next = (ring + sn * SL.floats)[SL.oS + SL.ezu + t];
            }
            const bool first = q == 0, last = q == v.ex - 1;
            const bool ym_ok = !(first && lo_y), yp_ok = !(last && hi_y), zm_ok = !(first && lo_z), zp_ok = !(last && hi_z);
            const float Sym = ym_ok ? cur[SL.oS + SL.ezu - v.ez] : 0.0f;
            const float Syp = yp_ok ? cur[SL.oS + SL.ezu + v.ez] : 0.0f;
            const float Szm = zm_ok ? cur[SL.oS + SL.ezu - 1] : 0.0f;
            const float Szp = zp_ok ? cur[SL.oS + SL.ezu + 1] : 0.0f;
            const float cyp = cur[SL.oCY + SL.ezu];
            const float cym = ym_ok ? cur[SL.oCY + SL.ezu - v.ez] : 0.0f;
            const float czp = cur[SL.oCZ + 4];
            const float czm = zm_ok ? cur[SL.oCZ + 3] : 0.0f;
            const float di = cur[SL.oDI];
            const float r = valid ? cur[SL.oR] : 0.0f;
            const float kv = v.kv ? cur[SL.oKV] : 0.0f;
            const int e15 = (int)(((int64_t)q * plane + m0) & 15);
            const uint8_t sd = reinterpret_cast<const uint8_t*>(cur - t + SL.oSD)[t + e15];
            __syncwarp();
            if (lane == 0) {   // this warp is done with the plane's stage (the next plane's stays for one more position)
                mbar_arrive_a(empty_a + 8u * st);
                if (j == T.xchunk) {   // last position of the task: nobody reads the following stage as "current"
                    mbar_wait_a(full_a + 8u * sn, (nn / kRing) & 1u);
                    mbar_arrive_a(empty_a + 8u * sn);
                }
            }
            float gd = di > 0.0f ? ((((((cxm + cxp) + cym) + cyp) + czm) + czp) + kv) * S0 : (di < 0.0f ? r : 0.0f);
            gd = fmaf(-cxm, S_prev, gd);
            gd = fmaf(-cxp, S_next, gd);
            gd = fmaf(-cym, Sym, gd);
            gd = fmaf(-cyp, Syp, gd);
            gd = fmaf(-czm, Szm, gd);
            gd = fmaf(-czp, Szp, gd);
            if (valid && (sd & 4)) {
                const int64_t e = (int64_t)q * plane + m;
                gd += v.G[e];
                v.G[e] = 0.0f;
            }
            S_prev = S0;
            cxm = cxp;
            const float g = valid ? gd * v.inv_n : 0.0f;
            const float x = __ldg(v.xe + q);
            const bool plus = (sd & 1) != 0;
            loss = fmaf(0.5f * r, r, loss);
            const bool do_p = plus && g != 0.0f;
            if (__any_sync(0xffffffffu, do_p)) P::grad_split(x, yz, do_p ? g : 0.0f, acc, par);
            if (!plus && g != 0.0f) {
                float a[NET::LMD][NET::HMW];
                M::template forward<P::NP>(x, y, z, a);
                M::template backward<P::NP, M::NP, P::NP>(x, y, z, a, g, accm);
            }
        }
        {   // the plane after the chunk was only read as "next": its position is skipped, its issue slot is not
            const unsigned ni = n_cur + kAhead;
            if (ni < n_total && ni % kWarps == warp) issue(ni);
            ++n_cur;
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float tj = (j & 1) ? hi32(acc.t[j / 2]) : lo32(acc.t[j / 2]);
            hs[j * kGradThreads] += tj;
            hs[(H + j) * kGradThreads] = fmaf(y, tj, hs[(H + j) * kGradThreads]);
            hs[(2 * H + j) * kGradThreads] = fmaf(z, tj, hs[(2 * H + j) * kGradThreads]);
        }
#pragma unroll
        for (int j = 0; j < HP2; ++j) acc.t[j] = 0ull;
    }
    loss *= v.inv_n;
#pragma unroll
    for (int i = 0; i <= NP; ++i) {
        float val;
        if (i < P::NP) val = P::split_get(acc, i, par, hs, kGradThreads);
        else if (i < NP) val = accm[i < NP ? i - P::NP : 0];
        else val = loss;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) red[warp * (NP + 1) + i] = val;
    }
    __syncthreads();
    for (int i = tid; i < NP + 1; i += kGradThreads) {
        float val = 0.0f;
#pragma unroll
        for (int w = 0; w < kGradThreads / 32; ++w) val += red[w * (NP + 1) + i];
        v.partials[(size_t)blockIdx.x * (NP + 1) + i] = val;
    }
}

template <class NET>
static cudaError_t launch_node_grad_fused(int grid, const FusedView& v, const Tasks& T, cudaStream_t st) {
    static int configured[64] = {0};   // per device
    const int bytes = fused_smem_bytes<NET>(v.ez);
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured[dev & 63] < bytes) {
        cudaError_t e = cudaFuncSetAttribute(node_grad_fused_kernel<NET>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = bytes;
    }
    node_grad_fused_kernel<NET><<<grid, kGradThreads, bytes, st>>>(v, T);
    return cudaSuccess;
}

#include "nbm_stencil_tma.cuh"

// ---------------------------------------------------------------------------------------------
// Learned preconditioner (nn/preconditioner.py:10-35; discretization.py:339, 418-419): every row is scaled by
// P = 0.5 + scale * sigmoid(MLP(coeffs_)), a tanh MLP 26 -> D1 -> D2 -> 1 on the point's 26 cell coefficients:
//   loss = mean 0.5 (P r)^2,  d loss/d r = P^2 r / n,  d loss/d theta_P = sum_p (P r^2 / n) dP/d theta_P.
// One pass over the lattice between the residual and the adjoint stage: it reads the un-preconditioned
// residual r = R[e], accumulates the loss and the preconditioner gradient, and overwrites R[e] <- P^2 r, which is
// what the adjoint/gradient stages then propagate.  The 26 x D1 outer product is accumulated co-operatively:
// a warp stages its 32 points' inputs and first-layer deltas in shared memory and lane k < 26 owns input row k.
// ---------------------------------------------------------------------------------------------
template <int D1, int D2>
struct PrecondNet {
    static constexpr int NIN = 26;
    static constexpr int oW1 = 0, ob1 = NIN * D1, oW2 = ob1 + D1, ob2 = oW2 + D1 * D2, oW3 = ob2 + D2, ob3 = oW3 + D2;
    static constexpr int NP = ob3 + 1;
};

__device__ __forceinline__ float tanh_acc(float x) { return tanh_nbm(x); }

template <int D1, int D2>
__global__ void __launch_bounds__(kThreads) precond_kernel(const float* __restrict__ coef26, int64_t cstride,
                                                           float* __restrict__ R, const int64_t* __restrict__ nodes,
                                                           int64_t ne, const float* __restrict__ params, float scale,
                                                           float inv_n, float* __restrict__ partials, int row_stride,
                                                           int col0, int loss_col) {
    using PN = PrecondNet<D1, D2>;
    constexpr int NIN = PN::NIN, NPc = PN::NP, kWarps = kThreads / 32;
    constexpr int CS = NIN + 1, DS = D1 + 1;   // padded strides: conflict-free column reads
    __shared__ float sP[NPc];
    __shared__ float sc[kWarps][32 * CS];
    __shared__ float sd[kWarps][32 * DS];
    __shared__ float sred[kWarps][NPc + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NPc; i += kThreads) sP[i] = params[i];
    __syncthreads();
    float aW1[D1];   // lane k < 26: d loss / d W1[k][:]
    float ab1[D1], aW2[D1][D2], ab2[D2], aW3[D2], ab3 = 0.0f, loss = 0.0f;
#pragma unroll
    for (int j = 0; j < D1; ++j) {
        aW1[j] = 0.0f;
        ab1[j] = 0.0f;
#pragma unroll
        for (int q = 0; q < D2; ++q) aW2[j][q] = 0.0f;
    }
#pragma unroll
    for (int q = 0; q < D2; ++q) { ab2[q] = 0.0f; aW3[q] = 0.0f; }
    for (int64_t base = (int64_t)blockIdx.x * kThreads; base < ne; base += (int64_t)gridDim.x * kThreads) {
        // `nodes` != null: the kernel walks a list of nodes (the crossed cells of the shared path) instead of a range
        const bool valid = base + tid < ne;
        const int64_t e = !valid ? 0 : (nodes ? nodes[base + tid] : base + tid);
        const float r = valid ? R[e] : 0.0f;
        float c[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) c[k] = valid ? __ldcs(coef26 + (int64_t)k * cstride + e) : 0.0f;
        // forward
        float h1[D1], h2[D2];
#pragma unroll
        for (int j = 0; j < D1; ++j) {
            float s = sP[PN::ob1 + j];
#pragma unroll
            for (int k = 0; k < NIN; ++k) s = fmaf(c[k], sP[PN::oW1 + k * D1 + j], s);
            h1[j] = tanh_acc(s);
        }
#pragma unroll
        for (int q = 0; q < D2; ++q) {
            float s = sP[PN::ob2 + q];
#pragma unroll
            for (int j = 0; j < D1; ++j) s = fmaf(h1[j], sP[PN::oW2 + j * D2 + q], s);
            h2[q] = tanh_acc(s);
        }
        float o = sP[PN::ob3];
#pragma unroll
        for (int q = 0; q < D2; ++q) o = fmaf(h2[q], sP[PN::oW3 + q], o);
        const float sg = 1.0f / (1.0f + __expf(-o));
        const float Pc = fmaf(scale, sg, 0.5f);
        const float pr = Pc * r;
        loss = fmaf(0.5f * pr, pr, loss);
        if (valid) R[e] = Pc * pr;                       // d loss / d r (times n)
        // backward: d loss/d P (times n) = P r^2
        const float dO = (pr * r) * inv_n * scale * sg * (1.0f - sg);
        ab3 += dO;
        float d2[D2], d1[D1];
#pragma unroll
        for (int q = 0; q < D2; ++q) {
            aW3[q] = fmaf(dO, h2[q], aW3[q]);
            d2[q] = dO * sP[PN::oW3 + q] * fmaf(-h2[q], h2[q], 1.0f);
            ab2[q] += d2[q];
        }
#pragma unroll
        for (int j = 0; j < D1; ++j) {
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < D2; ++q) {
                aW2[j][q] = fmaf(h1[j], d2[q], aW2[j][q]);
                s = fmaf(sP[PN::oW2 + j * D2 + q], d2[q], s);
            }
            d1[j] = s * fmaf(-h1[j], h1[j], 1.0f);
            ab1[j] += d1[j];
        }
        // co-operative outer product dW1[k][j] += c_p[k] d1_p[j] over the warp's 32 points
        if (__any_sync(0xffffffffu, r != 0.0f)) {
#pragma unroll
            for (int k = 0; k < NIN; ++k) sc[warp][lane * CS + k] = c[k];
#pragma unroll
            for (int j = 0; j < D1; ++j) sd[warp][lane * DS + j] = d1[j];
            __syncwarp();
            if (lane < NIN) {
                for (int pnt = 0; pnt < 32; ++pnt) {
                    const float ck = sc[warp][pnt * CS + lane];
#pragma unroll
                    for (int j = 0; j < D1; ++j) aW1[j] = fmaf(ck, sd[warp][pnt * DS + j], aW1[j]);
                }
            }
            __syncwarp();
        }
    }
    // block reduction -> partials[row][col0 + i], loss -> partials[row][loss_col]
#pragma unroll
    for (int j = 0; j < D1; ++j)
        if (lane < NIN) sred[warp][PN::oW1 + lane * D1 + j] = aW1[j];
    auto wsum = [&](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
#pragma unroll
    for (int j = 0; j < D1; ++j) {
        float v = wsum(ab1[j]);
        if (lane == 0) sred[warp][PN::ob1 + j] = v;
#pragma unroll
        for (int q = 0; q < D2; ++q) {
            v = wsum(aW2[j][q]);
            if (lane == 0) sred[warp][PN::oW2 + j * D2 + q] = v;
        }
    }
#pragma unroll
    for (int q = 0; q < D2; ++q) {
        float v = wsum(ab2[q]);
        if (lane == 0) sred[warp][PN::ob2 + q] = v;
        v = wsum(aW3[q]);
        if (lane == 0) sred[warp][PN::oW3 + q] = v;
    }
    {
        float v = wsum(ab3);
        if (lane == 0) sred[warp][PN::ob3] = v;
        v = wsum(loss);
        if (lane == 0) sred[warp][NPc] = v * inv_n;
    }
    __syncthreads();
    float* row = partials + (size_t)blockIdx.x * row_stride;
    for (int i = tid; i <= NPc; i += kThreads) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += sred[w][i];
        row[i < NPc ? col0 + i : loss_col] = v;
    }
}

// Uncrossed cells (all but ~0.5 %): the 26 inputs of a cell on side S are 6 face coefficients mu A / d of that side,
// the cell volume and the 6 face areas (grid constants), zeros for the other side (geometric_integrations_per_point.py:
// 906-996).  One launch per side with S a compile-time constant: the first layer is 6 x D1 FMAs on top of a per-side
// constant K, its gradient needs 6 x D1 accumulators plus sum(delta1) (the volume / area rows follow by scaling at the
// end), and every weight is a constant-bank operand.  Crossed cells take the generic kernel over their node list.
constexpr int kPcThreads = 384;   // 12 warps per SM at <= 168 registers

template <int D1, int D2, int S>
__global__ void __launch_bounds__(kPcThreads, 1) precond_bulk_kernel(const float* __restrict__ coef26, int64_t ne,
                                                                const int32_t* __restrict__ nodes, int64_t n_nodes,
                                                                float* __restrict__ R,
                                                                float vol, float ax, float ay, float az, float scale,
                                                                float inv_n, float* __restrict__ partials, int row_stride,
                                                                int col0, int loss_col) {
    using PN = PrecondNet<D1, D2>;
    constexpr int NPc = PN::NP, kWarps = kPcThreads / 32;
    __shared__ float sred[kWarps][NPc + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float area[6] = {ax, ax, ay, ay, az, az};
    // K[j] = b1[j] + vol W1[12+S][j] + sum_f A_f W1[14+2f+S][j]
    float K[D1];
#pragma unroll
    for (int j = 0; j < D1; ++j) {
        float k = fmaf(vol, c_PC[PN::oW1 + (12 + S) * D1 + j], c_PC[PN::ob1 + j]);
#pragma unroll
        for (int f = 0; f < 6; ++f) k = fmaf(area[f], c_PC[PN::oW1 + (14 + 2 * f + S) * D1 + j], k);
        K[j] = k;
    }
    // packed fp32x2 arithmetic, pairs along the output index (adjacent in the (in,out) row-major kernels)
    static_assert(D1 % 2 == 0 && D2 % 2 == 0, "packed preconditioner kernel needs even widths");
    constexpr int P1 = D1 / 2, P2 = D2 / 2;
    auto cp = [](int idx) { return *reinterpret_cast<const u64*>(&c_PC[idx]); };
    auto half = [](u64 v, int odd) { return odd ? hi32(v) : lo32(v); };
    const u64 k2l = pk(kTwoLog2e, kTwoLog2e), mone = pk(-1.0f, -1.0f);
    u64 K2[P1];
#pragma unroll
    for (int jp = 0; jp < P1; ++jp) K2[jp] = pk(K[2 * jp], K[2 * jp + 1]);
    u64 aW1[6][P1], aD[P1], aW2[D1][P2], ab2[P2], aW3[P2];
    float ab3 = 0.0f, loss = 0.0f;
#pragma unroll
    for (int jp = 0; jp < P1; ++jp) {
        aD[jp] = 0ull;
#pragma unroll
        for (int f = 0; f < 6; ++f) aW1[f][jp] = 0ull;
    }
#pragma unroll
    for (int j = 0; j < D1; ++j)
#pragma unroll
        for (int qp = 0; qp < P2; ++qp) aW2[j][qp] = 0ull;
#pragma unroll
    for (int qp = 0; qp < P2; ++qp) { ab2[qp] = 0ull; aW3[qp] = 0ull; }
    // the uncrossed row nodes of side S come as a list (built once per level); software pipeline (12 warps per SM cannot
    // hide dependent global loads): node index two entries ahead, its residual and 6 coefficients one entry ahead
    const int64_t stride = (int64_t)gridDim.x * kPcThreads;
    int64_t i = (int64_t)blockIdx.x * kPcThreads + tid;
    int e_0 = i < n_nodes ? __ldg(nodes + i) : -1;
    int e_1 = i + stride < n_nodes ? __ldg(nodes + i + stride) : -1;
    float r_n = 0.0f, c_n[6];
#pragma unroll
    for (int f = 0; f < 6; ++f) c_n[f] = 0.0f;
    if (e_0 >= 0) {
        r_n = R[e_0];
#pragma unroll
        for (int f = 0; f < 6; ++f) c_n[f] = __ldcs(coef26 + (int64_t)(2 * f + S) * ne + e_0);
    }
    for (; i < n_nodes; i += stride) {
        const int e = e_0;
        const float r = r_n;
        float c[6];
#pragma unroll
        for (int f = 0; f < 6; ++f) c[f] = c_n[f];
        e_0 = e_1;
        e_1 = i + 2 * stride < n_nodes ? __ldg(nodes + i + 2 * stride) : -1;
        if (e_0 >= 0) {
            r_n = R[e_0];
#pragma unroll
            for (int f = 0; f < 6; ++f) c_n[f] = __ldcs(coef26 + (int64_t)(2 * f + S) * ne + e_0);
        }
        if (r == 0.0f) continue;   // an exactly satisfied row contributes nothing
        u64 c2[6], h1[P1], h2[P2];
#pragma unroll
        for (int f = 0; f < 6; ++f) c2[f] = pk(c[f], c[f]);
#pragma unroll
        for (int jp = 0; jp < P1; ++jp) {
            u64 z = K2[jp];
#pragma unroll
            for (int f = 0; f < 6; ++f) z = ffma2(c2[f], cp(PN::oW1 + (2 * f + S) * D1 + 2 * jp), z);
            h1[jp] = tanh2_prescaled_4mufu(fmul2(z, k2l));
        }
#pragma unroll
        for (int qp = 0; qp < P2; ++qp) {
            u64 z = cp(PN::ob2 + 2 * qp);
#pragma unroll
            for (int j = 0; j < D1; ++j) {
                const float hj = half(h1[j / 2], j & 1);
                z = ffma2(pk(hj, hj), cp(PN::oW2 + j * D2 + 2 * qp), z);
            }
            h2[qp] = tanh2_prescaled_4mufu(fmul2(z, k2l));
        }
        u64 o2 = pk(c_PC[PN::ob3], 0.0f);
#pragma unroll
        for (int qp = 0; qp < P2; ++qp) o2 = ffma2(h2[qp], cp(PN::oW3 + 2 * qp), o2);
        const float o = lo32(o2) + hi32(o2);
        const float sg = 1.0f / (1.0f + __expf(-o));
        const float Pc = fmaf(scale, sg, 0.5f);
        const float pr = Pc * r;
        loss = fmaf(0.5f * pr, pr, loss);
        R[e] = Pc * pr;
        const float dO = (pr * r) * inv_n * scale * sg * (1.0f - sg);
        ab3 += dO;
        const u64 dO2 = pk(dO, dO), dO2n = pk(-dO, -dO);
        u64 d2[P2];
#pragma unroll
        for (int qp = 0; qp < P2; ++qp) {
            aW3[qp] = ffma2(dO2, h2[qp], aW3[qp]);
            const u64 omn = ffma2(h2[qp], h2[qp], mone);                       // h^2 - 1
            d2[qp] = fmul2(fmul2(cp(PN::oW3 + 2 * qp), dO2n), omn);            // dO W3 (1 - h^2)
            ab2[qp] = fadd2(ab2[qp], d2[qp]);
        }
        float tn[D1];   // -(W2 delta2)_j
#pragma unroll
        for (int j = 0; j < D1; ++j) {
            const float hj = half(h1[j / 2], j & 1);
            const u64 hj2 = pk(hj, hj);
            u64 t2 = 0ull;
#pragma unroll
            for (int qp = 0; qp < P2; ++qp) {
                aW2[j][qp] = ffma2(hj2, d2[qp], aW2[j][qp]);
                t2 = ffma2(cp(PN::oW2 + j * D2 + 2 * qp), d2[qp], t2);
            }
            tn[j] = -lo32(t2) - hi32(t2);
        }
#pragma unroll
        for (int jp = 0; jp < P1; ++jp) {
            const u64 omn = ffma2(h1[jp], h1[jp], mone);
            const u64 d1 = fmul2(pk(tn[2 * jp], tn[2 * jp + 1]), omn);        // (W2 delta2)(1 - h^2)
            aD[jp] = fadd2(aD[jp], d1);
#pragma unroll
            for (int f = 0; f < 6; ++f) aW1[f][jp] = ffma2(c2[f], d1, aW1[f][jp]);
        }
    }
    // block reduction into this CTA's partial row; rows of W1 that belong to the other side stay zero
    auto wsum = [&](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    for (int k = tid; k < kWarps * (NPc + 1); k += kPcThreads) (&sred[0][0])[k] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < D1; ++j) {
        float v = wsum(half(aD[j / 2], j & 1));
        if (lane == 0) {
            sred[warp][PN::ob1 + j] = v;
            sred[warp][PN::oW1 + (12 + S) * D1 + j] = vol * v;
#pragma unroll
            for (int f = 0; f < 6; ++f) sred[warp][PN::oW1 + (14 + 2 * f + S) * D1 + j] = area[f] * v;
        }
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            v = wsum(half(aW1[f][j / 2], j & 1));
            if (lane == 0) sred[warp][PN::oW1 + (2 * f + S) * D1 + j] = v;
        }
#pragma unroll
        for (int q = 0; q < D2; ++q) {
            v = wsum(half(aW2[j][q / 2], q & 1));
            if (lane == 0) sred[warp][PN::oW2 + j * D2 + q] = v;
        }
    }
#pragma unroll
    for (int q = 0; q < D2; ++q) {
        float v = wsum(half(ab2[q / 2], q & 1));
        if (lane == 0) sred[warp][PN::ob2 + q] = v;
        v = wsum(half(aW3[q / 2], q & 1));
        if (lane == 0) sred[warp][PN::oW3 + q] = v;
    }
    {
        float v = wsum(ab3);
        if (lane == 0) sred[warp][PN::ob3] = v;
        v = wsum(loss);
        if (lane == 0) sred[warp][NPc] = v * inv_n;
    }
    __syncthreads();
    float* row = partials + (size_t)blockIdx.x * row_stride;
    for (int i = tid; i <= NPc; i += kPcThreads) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += sred[w][i];
        row[i < NPc ? col0 + i : loss_col] = v;
    }
}

// forward only: Pc[i] = P(coeffs_ of point i)   (general path: the seeds of the backward pass need P first)
template <int D1, int D2>
__global__ void __launch_bounds__(kThreads) precond_fwd_kernel(const float* __restrict__ coef26, int64_t cstride,
                                                               int64_t n, const float* __restrict__ params, float scale,
                                                               float* __restrict__ Pc) {
    using PN = PrecondNet<D1, D2>;
    constexpr int NIN = PN::NIN, NPc = PN::NP;
    __shared__ float sP[NPc];
    for (int i = threadIdx.x; i < NPc; i += kThreads) sP[i] = params[i];
    __syncthreads();
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < n; e += (int64_t)gridDim.x * kThreads) {
        float c[NIN], h1[D1];
#pragma unroll
        for (int k = 0; k < NIN; ++k) c[k] = __ldg(coef26 + (int64_t)k * cstride + e);
#pragma unroll
        for (int j = 0; j < D1; ++j) {
            float s = sP[PN::ob1 + j];
#pragma unroll
            for (int k = 0; k < NIN; ++k) s = fmaf(c[k], sP[PN::oW1 + k * D1 + j], s);
            h1[j] = tanh_acc(s);
        }
        float o = sP[PN::ob3];
#pragma unroll
        for (int q = 0; q < D2; ++q) {
            float s = sP[PN::ob2 + q];
#pragma unroll
            for (int j = 0; j < D1; ++j) s = fmaf(h1[j], sP[PN::oW2 + j * D2 + q], s);
            o = fmaf(tanh_acc(s), sP[PN::oW3 + q], o);
        }
        Pc[e] = fmaf(scale, 1.0f / (1.0f + __expf(-o)), 0.5f);
    }
}

// K4a: deterministic sum of the per-CTA partial rows: one warp per column, lane l adds rows l, l + 32, ... in order,
// then a fixed shuffle tree (the row count and therefore the summation order depend only on the launch geometry)
// (`extra`: n_extra more addends of the LAST column - the per-CTA loss sums of the general path's rows kernel, kept as one
// compact strip instead of one mostly-zero partial row each)
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int rows, int np1, float* __restrict__ out,
                                       const float* __restrict__ extra = nullptr, int n_extra = 0) {
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= np1) return;
    float v = 0.0f;
    for (int r = lane; r < rows; r += 32) v += partials[(size_t)r * np1 + i];
    if (i == np1 - 1)
        for (int r = lane; r < n_extra; r += 32) v += extra[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[i] = v;
}

// K4b: optax chain (solvers/optimizers.py:33-54, optax 0.1.5 semantics), one CTA
__device__ __forceinline__ void optax_update(const nbm_optimizer_t& o, const float* __restrict__ loss_grad,
                                             float* __restrict__ params, float* __restrict__ state,
                                             int32_t* __restrict__ count, float* __restrict__ loss_hist) {
    __shared__ float red[32];
    __shared__ float s_scale;
    int P = o.n_params;
    float ss = 0.0f;
    for (int i = threadIdx.x; i < P; i += blockDim.x) ss = fmaf(loss_grad[i], loss_grad[i], ss);
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        float gn = sqrtf(t);
        // clip_by_global_norm: g if norm < max_norm else g / norm * max_norm
        s_scale = (o.optimizer == 0 && !(gn < o.max_norm)) ? (o.max_norm / gn) : 1.0f;
    }
    __syncthreads();
    int t_prev = *count;
    float t = (float)(t_prev + 1);
    float bc1 = 1.0f - powf(o.b1, t), bc2 = 1.0f - powf(o.b2, t);
    float step = o.lr;
    if (o.optimizer == 0) {
        if (o.scheduler == 0) step = o.lr * powf(o.decay_rate, (float)t_prev / o.transition_steps);
        else step = o.lr * (1.0f - fminf((float)t_prev, o.transition_steps) / o.transition_steps);
    }
    float* m = state;
    float* v = state + P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        float g = loss_grad[i] * s_scale;
        float upd;
        if (o.optimizer == 2) {  // optax.scale_by_rms: nu = d nu + (1-d) g^2 ; g * rsqrt(nu + eps)
            float vi = o.b2 * v[i] + (1.0f - o.b2) * (g * g);
            v[i] = vi;
            upd = g * rsqrtf(vi + o.eps);
        } else {
            float mi = (1.0f - o.b1) * g + o.b1 * m[i];
            float vi = (1.0f - o.b2) * (g * g) + o.b2 * v[i];
            m[i] = mi;
            v[i] = vi;
            upd = (mi / bc1) / (sqrtf(vi / bc2) + o.eps);
        }
        params[i] += -1.0f * (step * upd);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (loss_hist) loss_hist[t_prev] = loss_grad[P];
        *count = t_prev + 1;
    }
}

__global__ void apply_update_kernel(nbm_optimizer_t o, const float* __restrict__ loss_grad, float* __restrict__ params,
                                    float* __restrict__ state, int32_t* __restrict__ count,
                                    float* __restrict__ loss_hist) {
    optax_update(o, loss_grad, params, state, count, loss_hist);
}

// the three staged copies of network parameter i (see c_P): plain, pre-scaled by 2 log2(e), transposed + negated
__device__ __forceinline__ void stage_param(const nbm_net_t& net, int i, float v, float* __restrict__ stage) {
    int np = 3 * net.hidden_p + net.hidden_p + (net.layers_p - 1) * (net.hidden_p * net.hidden_p + net.hidden_p) +
             net.hidden_p + 1;
    int k = i < np ? i : i - np;
    int H = i < np ? net.hidden_p : net.hidden_m;
    int Lh = i < np ? net.layers_p : net.layers_m;
    int hidden_len = 4 * H + (Lh - 1) * (H * H + H);  // everything before the output layer
    stage[i] = v;
    stage[NBM_MAXP + i] = k < hidden_len ? v * kTwoLog2e : v;
    // third copy: hidden HxH matrices transposed in place AND negated: the backward pass forms (a^2 - 1) = -(1 - a^2)
    // with one FFMA2 (no sign flips) and gets the sign back from -W^T
    int dst = i;
    float v3 = v;
    if (k >= 4 * H && k < hidden_len) {
        int q = (k - 4 * H) % (H * H + H);
        if (q < H * H) {
            dst = i - q + (q % H) * H + q / H;
            v3 = -v;
        }
    }
    stage[2 * NBM_MAXP + dst] = v3;
}

// K4 fused: [partial rows -> loss_grad] -> optax chain -> staged parameter copies for the next step.  One CTA of 1024.
__global__ void __launch_bounds__(1024) finalize_step_kernel(nbm_optimizer_t o, nbm_net_t net, int n_net,
                                                             const float* __restrict__ partials, int rows, int np1,
                                                             float* __restrict__ loss_grad, float* __restrict__ params,
                                                             float* __restrict__ state, int32_t* __restrict__ count,
                                                             float* __restrict__ loss_hist, float* __restrict__ stage) {
    extern __shared__ float sred[];   // [ngrp][np1]
    const int t = threadIdx.x;
    pdl_wait();
    if (partials) {
        const int ngrp = max(1, (int)blockDim.x / np1);
        const int grp = t / np1, col = t - grp * np1;
        if (grp < ngrp) {
            float v = 0.0f;
            for (int r = grp; r < rows; r += ngrp) v += partials[(size_t)r * np1 + col];
            sred[grp * np1 + col] = v;
        }
        __syncthreads();
        if (t < np1) {
            float v = 0.0f;
            for (int g = 0; g < ngrp; ++g) v += sred[g * np1 + t];
            loss_grad[t] = v;
        }
        __syncthreads();
    }
    optax_update(o, loss_grad, params, state, count, loss_hist);
    __syncthreads();
    for (int i = t; i < n_net; i += blockDim.x) stage_param(net, i, params[i], stage);
}

// K5: evaluation (trainer.py:960-977)
template <class NET, int LP, int HP, int LM, int HM>
__global__ void evaluate_kernel(nbm_lvl_t L, const float* __restrict__ pts, int64_t n, float dx, float dy, float dz,
                                float* __restrict__ u, float* __restrict__ grad_u, float* __restrict__ grad_n) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float x = pts[3 * e], y = pts[3 * e + 1], z = pts[3 * e + 2];
    const float* ep = L.eval_phi ? L.eval_phi + 7 * e : nullptr;   // sampled level set (see nbm_lvl_t)
    float ph = ep ? ep[0] : phi_at(L, x, y, z);
    float gx[3], val;
    if (ph >= 0.0f) {
        float a[LP][HP];
        val = Mlp<LP, HP>::template forward<0>(x, y, z, a);
        Mlp<LP, HP>::template input_grad<0>(a, gx);
    } else {
        float a[LM][HM];
        val = Mlp<LM, HM>::template forward<Mlp<LP, HP>::NP>(x, y, z, a);
        Mlp<LM, HM>::template input_grad<Mlp<LP, HP>::NP>(a, gx);
    }
    u[e] = val;
    if (grad_u) { grad_u[3 * e] = gx[0]; grad_u[3 * e + 1] = gx[1]; grad_u[3 * e + 2] = gx[2]; }
    if (grad_n) {
        float px, py, pz;
        if (ep) {
            px = (ep[2] - ep[1]) / (2.0f * dx);
            py = (ep[4] - ep[3]) / (2.0f * dy);
            pz = (ep[6] - ep[5]) / (2.0f * dz);
        } else {
            px = (phi_at(L, x + dx, y, z) - phi_at(L, x - dx, y, z)) / (2.0f * dx);
            py = (phi_at(L, x, y + dy, z) - phi_at(L, x, y - dy, z)) / (2.0f * dy);
            pz = (phi_at(L, x, y, z + dz) - phi_at(L, x, y, z - dz)) / (2.0f * dz);
        }
        float nrm = sqrtf(px * px + py * py + pz * pz);
        grad_n[e] = (px / nrm) * gx[0] + (py / nrm) * gx[1] + (pz / nrm) * gx[2];
    }
}


// staging copy of the parameters: [0,MAXP) as given, [MAXP,2MAXP) with the hidden layers pre-scaled
__global__ void prep_params_kernel(nbm_net_t net, const float* __restrict__ params, float* __restrict__ stage, int P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    stage_param(net, i, params[i], stage);
}


// ---------------------------------------------------------------------------------------------
// fused partial-row reduction + all-reduce over NVLink peer memory (one process per GPU, CUDA IPC)
// ---------------------------------------------------------------------------------------------
struct CommBlock {
    unsigned int flag[2][NBM_COMM_MAX_RANKS];   // [parity][source rank]: step + 1 of the last slot that rank pushed here
    unsigned int error;                         // set when a peer wait timed out
    unsigned int pad[3];
    float data[2][NBM_COMM_MAX_RANKS][NBM_MAXP + 8];   // [parity][source rank][entry], written by the source rank
};

struct CommPeers {
    CommBlock* b[NBM_COMM_MAX_RANKS];
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// One CTA.  (1) deterministic sum of the partial rows (row groups in parallel, fixed order);  (2) PUSH: every rank stores
// its [grad, loss] vector into its slot of EVERY rank's block (remote stores over NVLink are fire-and-forget), fences,
// then raises its flag in every block;  (3) each rank polls only its LOCAL flags and adds the slots in rank order, so the
// result is bitwise identical on all ranks.  Slots are double-buffered by step parity: a peer can be at most one step
// ahead (it cannot finish step k+1 before this rank has raised its k+1 flags, which happens after step k's reads).
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// `timeout_ns` == 0: wait for the peers without a bound (what NCCL does).  Otherwise a rank that has waited that long
// sets the block's error word and poisons ITS output with NaN - never a plausible-looking partial sum; the host turns
// the error word into an exception (Trainer._check_comm, bench.py).
__device__ __forceinline__ void reduce_allreduce_body(const float* __restrict__ partials, int rows, int np1,
                                                                int rank, int world, CommPeers peers,
                                                                int32_t* __restrict__ step_dev, float* __restrict__ out,
                                                                unsigned long long timeout_ns) {
    extern __shared__ float sred[];   // [ngrp][np1]
    const int t = threadIdx.x;
    pdl_wait();
    const unsigned int step = (unsigned int)*step_dev;
    const int par = step & 1;
    CommBlock* mine = peers.b[rank];
    const int ngrp = max(1, (int)blockDim.x / np1);
    {
        const int grp = t / np1, col = t - grp * np1;
        if (grp < ngrp) {
            float v = 0.0f;
            for (int r = grp; r < rows; r += ngrp) v += partials[(size_t)r * np1 + col];
            sred[grp * np1 + col] = v;
        }
    }
    __syncthreads();
    if (t < np1) {
        float v = 0.0f;
        for (int g = 0; g < ngrp; ++g) v += sred[g * np1 + t];
        for (int r = 0; r < world; ++r) peers.b[r]->data[par][rank][t] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_fail;
    if (t == 0) s_fail = 0;
    if (t < world) st_release_sys(&peers.b[t]->flag[par][rank], step + 1);
    __syncthreads();
    if (t < world && t != rank) {
        const unsigned int* f = &mine->flag[par][t];
        const unsigned long long t0 = timeout_ns ? global_ns() : 0ull;
        unsigned int spins = 0;
        while ((int)(ld_acquire_sys(f) - (step + 1)) < 0) {
            if (timeout_ns && ((++spins & 1023u) == 0u) && global_ns() - t0 > timeout_ns) {
                atomicExch(&mine->error, 1u);
                atomicExch(&s_fail, 1);
                break;
            }
        }
    }
    __syncthreads();
    if (t < np1) {
        float v = 0.0f;
        for (int r = 0; r < world; ++r) v += ld_relaxed_sys(&mine->data[par][r][t]);
        out[t] = s_fail ? __int_as_float(0x7fc00000) : v;
    }
    __syncthreads();
    if (t == 0) *step_dev = (int32_t)(step + 1);
}

__global__ void __launch_bounds__(1024) reduce_allreduce_kernel(const float* __restrict__ partials, int rows, int np1,
                                                                int rank, int world, CommPeers peers,
                                                                int32_t* __restrict__ step_dev, float* __restrict__ out,
                                                                unsigned long long timeout_ns) {
    reduce_allreduce_body(partials, rows, np1, rank, world, peers, step_dev, out, timeout_ns);
}

// K4 on several GPUs as ONE kernel: partial rows -> exchange over peer memory -> optax chain -> staged parameter copies
// (reduce_allreduce_kernel followed by finalize_step_kernel's tail: one launch and one kernel boundary less per step)
__global__ void __launch_bounds__(1024) reduce_allreduce_finalize_kernel(
    const float* __restrict__ partials, int rows, int np1, int rank, int world, CommPeers peers, int32_t* __restrict__ step_dev,
    float* __restrict__ loss_grad, unsigned long long timeout_ns, nbm_optimizer_t o, nbm_net_t net, int n_net,
    float* __restrict__ params, float* __restrict__ state, int32_t* __restrict__ count, float* __restrict__ loss_hist,
    float* __restrict__ stage) {
    reduce_allreduce_body(partials, rows, np1, rank, world, peers, step_dev, loss_grad, timeout_ns);
    __syncthreads();
    optax_update(o, loss_grad, params, state, count, loss_hist);
    __syncthreads();
    for (int i = threadIdx.x; i < n_net; i += blockDim.x) stage_param(net, i, params[i], stage);
}

static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

bool pdl_enabled() {
    // opt-in (NBM_PDL=1).  Measured on B200 with the step as a CUDA graph: 256^3 625.7 us with, 622.8 without;
    // 128^3 119.7 / 120.6; 64^3 47.3 / 47.7 - the graph's kernel-to-kernel edges are already that cheap, and CTAs of
    // the next kernel placed during the last wave of the previous one take its resources.
    static const bool on = getenv("NBM_PDL") && getenv("NBM_PDL")[0] == '1';
    return on;
}

// library-owned side stream + fork/join events of the list chain, one set per device (created on first use)
struct SideLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static SideLane* side_lane() {
    static SideLane lanes[64];
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    SideLane& l = lanes[dev & 63];
    std::lock_guard<std::mutex> lock(mu);
    if (!l.stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // (hi = numerically lowest = highest priority)
        const char* pe = getenv("NBM_SIDE_PRIO");     // timing experiments: "low" = lowest priority
        if (cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, (pe && pe[0] == 'l') ? lo : hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) {
            l.stream = nullptr;
            return nullptr;
        }
    }
    return &l;
}

constexpr int kPartialRows = 148 * 9;  // shared path: <= 148 rows; general path: 7 replicas x 148 + rows + extrap rows

template <class NET>
static int launch_shared(const nbm_shared_step_t& s, cudaStream_t st) {
    const int sms = sm_count();
    const int stages = s.stages == 0 ? 0x3f : s.stages;
    // x-chunk length: 16 planes on large lattices (amortises the task set-up and the exposed first load); shorter on
    // small ones so that every SM still gets >= 2 tasks
    int xchunk = 16;
    {
        int mblocks = (s.ey * s.ez + kThreads - 1) / kThreads;
        while (xchunk > 2 && (int64_t)mblocks * ((s.ex + xchunk - 1) / xchunk) < 2 * (int64_t)sms) xchunk >>= 1;
    }
    Tasks T = make_tasks(s.ex, s.ey, s.ez, xchunk);
    // 16-byte paths need aligned rows: plane % 4 == 0, ez even, and 16-byte aligned array bases
    const bool vec4 = ((s.ey * s.ez) % 4 == 0) && (s.ez % 2 == 0) &&
                      ((((uintptr_t)s.w | (uintptr_t)s.rhs | (uintptr_t)s.U | (uintptr_t)s.R | (uintptr_t)s.G |
                         (uintptr_t)s.nl) & 15) == 0);
    // the gradient kernel runs one 384-thread CTA per SM: its own strips, x chunks sized for >= 2 tasks per CTA
    int xchunk_g = 16;
    {
        int mblocks = (s.ey * s.ez + kGradThreads - 1) / kGradThreads;
        while (xchunk_g > 2 && (int64_t)mblocks * ((s.ex + xchunk_g - 1) / xchunk_g) < 2 * (int64_t)sms) xchunk_g >>= 1;
    }
    Tasks Tg = make_tasks(s.ex, s.ey, s.ez, xchunk_g, kGradThreads);
    const bool pc = s.coef26 != nullptr;
    const int n_pc = pc ? PrecondNet<8, 4>::NP : 0;
    const int pc_stride = NET::NP + n_pc + 1;
    const int gridP = pc ? min(s.n_pc_rows, 3 * sms) : 0;
    int gridC = (int)min((int64_t)Tg.mblocks * s.ex, (int64_t)min(kPartialRows, sms));
    if (gridC > s.n_partial_rows - gridP) gridC = s.n_partial_rows - gridP;
    Tg.split = run_split((int64_t)Tg.mblocks * s.ex, gridC);
    // K_B: residual rows + adjoint stencil of the faces table as one TMA-fed kernel whenever both stages are wanted and
    // nothing sits between them (the preconditioner rescales R; the r1 fused gradient kernel wants S)
    const int dense2 = NBM_STAGE_RESIDUAL | NBM_STAGE_ADJOINT;
    // (not with a nonlinear operator: sinh / cosh per cell on the one-barrier-per-plane critical path of the TMA kernel
    // measured 510 us at Poisson-Boltzmann 256^3 against 127 + 131 us for the two separate kernels)
    const bool fusedA = s.stencil_tma >= 0 && s.faces && !pc && !s.S && !s.nl && ((stages & dense2) == dense2) &&
                        stencil_tma::applicable(s);
    // the two separate face-table kernels can take the list chain beside them just as well
    const bool sepA = !fusedA && s.faces && !pc && !s.S && ((stages & dense2) == dense2);
    // timing modifiers: run only the dense kernels / only the list kernels of the selected stages
    const bool lists_on = !(stages & NBM_STAGE_NO_LISTS), dense_on = !(stages & NBM_STAGE_NO_DENSE);
    if (dense_on && (stages & NBM_STAGE_FWD)) {
        // 3 CTAs are resident per SM.  Large lattices: 12 CTAs per SM (4 even waves; measured 153.6 us at 256^3 against
        // 154.6 with 6 and 160.9 with 3: minus-side regions make CTAs uneven).  Small ones: fewer, so that a CTA keeps
        // >= 24 plane iterations per range start, down to one wave.
        static const int fwd_env = getenv("NBM_FWD_CTAS") ? atoi(getenv("NBM_FWD_CTAS")) : 0;
        const int64_t iters = (int64_t)T.mblocks * s.ex;
        int per_sm = fwd_env > 0 ? fwd_env : 12;
        while (!fwd_env && per_sm > 3 && iters / ((int64_t)sms * per_sm) < 24) per_sm -= 3;
        int gridA = (int)min(iters, (int64_t)sms * per_sm);
        if (s.Hst) fwd_nodes_kernel<NET, false, true><<<gridA, kThreads, 0, st>>>(view_of(s), T);
        else fwd_nodes_kernel<NET, false, false><<<gridA, kThreads, 0, st>>>(view_of(s), T);
    }
    // the list chain beside the dense stencil (see nbm_shared_step_t.G2): whole-step launches only
    const int chain = NBM_STAGE_EXTRAP | dense2;
    const bool overlap = (fusedA || sepA) && lists_on && dense_on && ((stages & chain) == chain) && s.G2 && s.Rq && s.list_nodes &&
                         !s.g_ptr && (s.n_irr > 0 || s.n_crossed > 0) && (s.n_irr == 0 || s.n_list > 0);
    if (overlap) {
        SideLane* lane = side_lane();
        if (!lane) return cuda_check(cudaGetLastError(), "side stream of the list chain");
        static const int dbg_skip = getenv("NBM_OV_SKIP") ? atoi(getenv("NBM_OV_SKIP")) : 0;   // timing experiments only
        cudaEventRecord(lane->fork, st);
        cudaStreamWaitEvent(lane->stream, lane->fork, 0);
        if (!(dbg_skip & 1)) {
        if (s.n_crossed > 0) extrap_kernel<<<(unsigned)((s.n_crossed + 127) / 128), 128, 0, lane->stream>>>(s);
        if (s.n_irr > 0)
            launch_pdl(irregular_fb_kernel, dim3((unsigned)((s.n_irr + 127) / 128)), dim3(128), 0, lane->stream, s);
        if (s.n_crossed > 0)
            launch_pdl(extrap_bwd_kernel, dim3((unsigned)((s.n_crossed * 27 + 127) / 128)), dim3(128), 0, lane->stream, s, s.G2);
        }
        cudaEventRecord(lane->join, lane->stream);
        if (fusedA) {
            int rc = stencil_tma::launch(s, sms, st);
            if (rc) return rc;
        } else {   // (irregular rows keep R = 0 until the merge: their nonlinear centre term is irregular_fb's)
            const unsigned gx = (unsigned)((s.ey * s.ez / 4 + kThreads - 1) / kThreads);
            residual_faces4_kernel<<<dim3(gx, s.ex - 2), kThreads, 0, st>>>(s);
            adjoint_faces4_kernel<<<dim3(gx, s.ex), kThreads, 0, st>>>(s);
        }
        cudaStreamWaitEvent(st, lane->join, 0);
        const int64_t nm = s.n_list > s.n_irr ? s.n_list : s.n_irr;
        if (!(dbg_skip & 2)) merge_lists_kernel<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(s);
    }
    if (!overlap && (stages & NBM_STAGE_EXTRAP) && lists_on && s.n_crossed > 0)
        extrap_kernel<<<(unsigned)((s.n_crossed + 127) / 128), 128, 0, st>>>(s);
    if (overlap) {
    } else if (fusedA) {
        if (dense_on) {
            int rc = stencil_tma::launch(s, sms, st);
            if (rc) return rc;
        }
        // (the dense adjoint ran before these rows' residuals exist: irregular_bwd adds their nonlinear centre term)
        if (lists_on && s.n_irr > 0) irregular_fwd_kernel<<<(unsigned)((s.n_irr + 127) / 128), 128, 0, st>>>(s);
    } else if (stages & NBM_STAGE_RESIDUAL) {
        if (!dense_on) {
        } else if (s.faces) {
            dim3 g((s.ey * s.ez / 4 + kThreads - 1) / kThreads, s.ex - 2);
            residual_faces4_kernel<<<g, kThreads, 0, st>>>(s);
        } else if (vec4) {
            dim3 g((s.ey * s.ez / 4 + kThreads - 1) / kThreads, s.ex - 2);
            residual4_kernel<<<g, kThreads, 0, st>>>(s);
        } else {
            dim3 g((s.ey * s.ez + kThreads - 1) / kThreads, s.ex - 2);
            residual_kernel<<<g, kThreads, 0, st>>>(s);
        }
        if (lists_on && s.n_irr > 0) irregular_fwd_kernel<<<(unsigned)((s.n_irr + 127) / 128), 128, 0, st>>>(s);
        if (pc) {
            // rows [gridC, gridC + gridP) of the partials: preconditioner gradient + the loss.  Uncrossed cells: one
            // launch per side (inputs reduce to 6 face coefficients); crossed cells: the generic kernel on their list
            const int64_t ne = (int64_t)s.ex * s.ey * s.ez;
            cudaMemcpyToSymbolAsync(c_PC, s.pc_params, sizeof(float) * n_pc, 0, cudaMemcpyDeviceToDevice, st);
            float* rows = s.partials + (size_t)gridC * pc_stride;
            if (s.pc_nodes_m || s.pc_nodes_p) {
                const float vol = s.pc_d[0] * s.pc_d[1] * s.pc_d[2];
                const float ax = s.pc_d[1] * s.pc_d[2], ay = s.pc_d[0] * s.pc_d[2], az = s.pc_d[0] * s.pc_d[1];
                const int gb = gridP / 3, gx = gridP - 2 * gb;
                // (all three always launched: each also zero-fills its partial rows)
                precond_bulk_kernel<8, 4, 0><<<gb, kPcThreads, 0, st>>>(s.coef26, ne, s.pc_nodes_m, s.n_pc_m, s.R, vol, ax, ay, az,
                                                                     s.pc_scale, s.inv_n_points, rows, pc_stride, NET::NP,
                                                                     NET::NP + n_pc);
                precond_bulk_kernel<8, 4, 1><<<gb, kPcThreads, 0, st>>>(s.coef26, ne, s.pc_nodes_p, s.n_pc_p, s.R, vol, ax, ay, az,
                                                                     s.pc_scale, s.inv_n_points, rows + (size_t)gb * pc_stride,
                                                                     pc_stride, NET::NP, NET::NP + n_pc);
                precond_kernel<8, 4><<<gx, kThreads, 0, st>>>(s.coef26, ne, s.R, s.c_node, s.n_crossed, s.pc_params, s.pc_scale,
                                                              s.inv_n_points, rows + (size_t)2 * gb * pc_stride, pc_stride,
                                                              NET::NP, NET::NP + n_pc);
            } else {
                precond_kernel<8, 4><<<gridP, kThreads, 0, st>>>(s.coef26, ne, s.R, nullptr, ne, s.pc_params, s.pc_scale,
                                                                 s.inv_n_points, rows, pc_stride, NET::NP, NET::NP + n_pc);
            }
        }
    }
    const bool fused = s.faces && s.S && !s.nl && !pc;
    if ((stages & NBM_STAGE_ADJOINT) && !overlap) {
        if (fusedA || !dense_on) {
            // (dense adjoint already done inside K_B / not wanted)
        } else if (fused) {
            // the dense adjoint stencil runs inside the gradient kernel; G only collects the list contributions.
            // A profiling run that stops before the gradient stage would leave them behind: clear first.
            if (!(stages & NBM_STAGE_GRAD))
                cudaMemsetAsync(s.G, 0, sizeof(float) * (size_t)s.ex * s.ey * s.ez, st);
        } else if (s.faces) {
            dim3 g((s.ey * s.ez / 4 + kThreads - 1) / kThreads, s.ex);
            adjoint_faces4_kernel<<<g, kThreads, 0, st>>>(s);
        } else if (vec4) {
            dim3 g((s.ey * s.ez / 4 + kThreads - 1) / kThreads, s.ex);
            adjoint4_kernel<<<g, kThreads, 0, st>>>(s);
        } else {
            dim3 g((s.ey * s.ez + kThreads - 1) / kThreads, s.ex);
            adjoint_kernel<<<g, kThreads, 0, st>>>(s);
        }
        if (!lists_on) {
        } else if (s.g_ptr) {   // deterministic gathers through the transposed incidence
            if (s.n_crossed > 0) gather_gE_kernel<<<(unsigned)((s.n_crossed + 127) / 128), 128, 0, st>>>(s);
            if (s.n_list > 0) gather_G_kernel<<<(unsigned)((s.n_list + 127) / 128), 128, 0, st>>>(s);
        } else {
            if (s.n_irr > 0) irregular_bwd_kernel<<<(unsigned)((s.n_irr + 127) / 128), 128, 0, st>>>(s, fusedA);
            if (s.n_crossed > 0) extrap_bwd_kernel<<<(unsigned)((s.n_crossed * 27 + 127) / 128), 128, 0, st>>>(s, s.G);
        }
    }
    if ((stages & NBM_STAGE_GRAD) && fused) {
        FusedView f;
        f.xe = s.xe; f.ye = s.ye; f.ze = s.ze;
        f.ex = s.ex; f.ey = s.ey; f.ez = s.ez;
        f.ne = (int64_t)s.ex * s.ey * s.ez;
        f.side = s.side; f.S = s.S; f.cface = s.cface; f.dinv = s.dinv; f.kv = s.kv; f.R = s.R; f.G = s.G;
        f.inv_n = s.inv_n_points; f.partials = s.partials;
        cudaError_t e = launch_node_grad_fused<NET>(gridC, f, Tg, st);
        if (e != cudaSuccess) return cuda_check(e, "node_grad_fused attribute");
    } else if (stages & NBM_STAGE_GRAD) {
        NodeView nv = view_of(s);
        if (pc) {   // the loss comes from the preconditioner kernel; rows carry the preconditioner's columns too
            nv.R = nullptr;
            nv.row_stride = pc_stride;
            nv.loss_col = NET::NP + n_pc;
        }
        // (a call without the forward stage reads the stash of the last forward, as it reads that forward's U)
        const bool stash = nv.Hst != nullptr;
        cudaError_t e = stash ? launch_node_grad<NET, false, true>(dim3(gridC), nv, Tg, st)
                              : launch_node_grad<NET, false, false>(dim3(gridC), nv, Tg, st);
        if (e != cudaSuccess) return cuda_check(e, "node_grad attribute");
    }
    if (stages & NBM_STAGE_REDUCE)
        reduce_partials_kernel<<<(pc_stride * 32 + 127) / 128, 128, 0, st>>>(s.partials, gridC + gridP, pc_stride, s.loss_grad);
    return cuda_check(cudaGetLastError(), "shared step launch");
}

// FP32 FMA-pipe probe: 16 independent accumulator chains per thread
__global__ void __launch_bounds__(256) ffma_probe_kernel(int iters, float* out) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i);
    float b = 1.0000001f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) out[1] = s;  // keeps the chains alive without memory traffic
}

#define NBM_NET_DISPATCH(net, FN, ...)                                                                       \
    do {                                                                                                     \
        const nbm_net_t& _n = (net);                                                                         \
        if (_n.layers_p == 2 && _n.hidden_p == 10 && _n.layers_m == 1 && _n.hidden_m == 1)                   \
            return FN<Net<2, 10, 1, 1>>(__VA_ARGS__);                                                        \
        if (_n.layers_p == 2 && _n.hidden_p == 10 && _n.layers_m == 1 && _n.hidden_m == 3)                   \
            return FN<Net<2, 10, 1, 3>>(__VA_ARGS__);                                                        \
        if (_n.layers_p == 1 && _n.hidden_p == 10 && _n.layers_m == 1 && _n.hidden_m == 1)                   \
            return FN<Net<1, 10, 1, 1>>(__VA_ARGS__);                                                        \
        if (_n.layers_p == 3 && _n.hidden_p == 10 && _n.layers_m == 1 && _n.hidden_m == 1)                   \
            return FN<Net<3, 10, 1, 1>>(__VA_ARGS__);                                                        \
        set_error("network shape p(%d x %d) m(%d x %d) is outside the compiled kernel set", _n.layers_p,     \
                  _n.hidden_p, _n.layers_m, _n.hidden_m);                                                    \
        return NBM_ERR_UNSUPPORTED;                                                                          \
    } while (0)

static int dispatch_shared(const nbm_shared_step_t& s, cudaStream_t st) { NBM_NET_DISPATCH(s.net, launch_shared, s, st); }


// =============================================================================================
// General path: any cell size (multi-resolution levels, data_management.py:320-326) and any
// contiguous batch.  Stencil sites p +- d e_a are not grid nodes, nothing is shared between points:
// 7 network evaluations per point (the reference does 197), fused forward + residual + backward.
// =============================================================================================
struct PointsArgs {
    nbm_points_step_t s;
    float shift[7][3];
    int64_t n_points;  // nx*ny*nz
};

// Z0: E[c] for the crossed sites of the batch: one warp per site, lane q < 27 evaluates the cube
// vertex s + X_q (get_Xijk, discretization.py:164-197: x fastest)
template <class NET>
__global__ void __launch_bounds__(kThreads) points_extrap_kernel(PointsArgs a) {
    const nbm_points_step_t& s = a.s;
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < s.n_crossed; c += nwarps) {
        int64_t p = s.c_site[c] % a.n_points;
        if (p < s.p0 || p >= s.p1) continue;
        float v = 0.0f;
        if (lane < 27) {
            float X0 = (float)(lane % 3 - 1) * s.dx, X1 = (float)((lane / 3) % 3 - 1) * s.dy,
                  X2 = (float)(lane / 9 - 1) * s.dz;
            bool plus = (s.c_cube_side[c] >> lane) & 1u;
            float u = NET::eval(plus, s.c_pos[3 * c] + X0, s.c_pos[3 * c + 1] + X1, s.c_pos[3 * c + 2] + X2);
            v = s.B[c * 28 + lane] * u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            s.E[c] = v + s.B[c * 28 + 27];
            s.gE[c] = 0.0f;
        }
    }
}

// Z1: rows of the batch, pointwise on the 7 site values U7[k][p] (written by fwd_nodes over the 7 displaced
// lattices): residual, loss partial, and d loss / d u(site) into G7[k][p] for the per-site backward kernel
// S4: the 7 site values come from the 4 shared lattices of zoom level 1 (nbm_points_step_t.U4), padded dims
// (nx+1, ny+1, nz+1); a half-offset site belongs to two points, whose contributions meet in G4 by atomicAdd (the forward
// kernel cleared it; two addends: the order cannot change the sum)
template <bool S4>
__global__ void __launch_bounds__(kThreads) points_rows_kernel(PointsArgs a, const float* __restrict__ U7,
                                                               float* __restrict__ G7, int row0, int np1,
                                                               float* __restrict__ loss_strip) {
    const nbm_points_step_t& s = a.s;
    const int64_t N = a.n_points;
    const int64_t sy4 = s.nz + 1, sx4 = (int64_t)(s.ny + 1) * sy4, ne4 = (int64_t)(s.nx + 1) * sx4;
    float loss = 0.0f;
    for (int64_t p = s.p0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < s.p1;
         p += (int64_t)gridDim.x * blockDim.x) {
        float u[7], w[7];
        float r = -__ldg(s.rhs + p);
        int64_t i4[7];
        if (S4) {
            const int64_t pl = (int64_t)s.ny * s.nz;
            const int ix = (int)(p / pl), rem = (int)(p - (int64_t)ix * pl), iy = rem / s.nz, iz = rem - iy * s.nz;
            const int64_t e4 = ix * sx4 + iy * sy4 + iz;
            i4[0] = e4;
            i4[1] = ne4 + e4;     i4[2] = ne4 + e4 + sx4;
            i4[3] = 2 * ne4 + e4; i4[4] = 2 * ne4 + e4 + sy4;
            i4[5] = 3 * ne4 + e4; i4[6] = 3 * ne4 + e4 + 1;
        }
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            w[q] = __ldg(s.w + q * N + p);
            u[q] = S4 ? U7[i4[q]] : U7[q * N + p];
            r = fmaf(w[q], u[q], r);
        }
        float nlw0 = 0.0f, nlw1 = 0.0f;
        if (s.nl) {
            nlw0 = s.nl[p];
            nlw1 = s.nl[N + p];
            r = fmaf(nlw0, nl_apply(s.nonlinear_m, s.nl_coef_m, u[0]), r);
            r = fmaf(nlw1, nl_apply(s.nonlinear_p, s.nl_coef_p, u[0]), r);
        }
        int32_t q_irr = s.irr[p];
        if (q_irr >= 0) {
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                int32_t c = s.irr_c[(int64_t)q_irr * 7 + q];
                if (c >= 0) r = fmaf(s.irr_wE[(int64_t)q_irr * 7 + q], s.E[c], r);
            }
            uint8_t nlr = s.irr_nl[q_irr];
            if (nlr) {
                float Ec = s.E[s.irr_c[(int64_t)q_irr * 7]];
                r = fmaf(s.irr_nlw[q_irr], nlr == 1 ? nl_apply(s.nonlinear_m, s.nl_coef_m, Ec)
                                                    : nl_apply(s.nonlinear_p, s.nl_coef_p, Ec), r);
            }
        }
        if (s.rows) s.rows[p] = r;
        if (s.Pc) {
            // preconditioned row P r: d loss/d r (times n) = P^2 r; loss and d/d theta_P come from precond_kernel
            const float Pp = s.Pc[p];
            r *= Pp * Pp;
        } else {
            loss = fmaf(0.5f * r, r, loss);
        }
        if (q_irr >= 0) {
            // each crossed site belongs to exactly one (point, slot): plain stores (gE carries the un-normalised
            // residual weight; the 1/n of the mean is applied by the backward kernels)
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                int32_t c = s.irr_c[(int64_t)q_irr * 7 + q];
                if (c >= 0) {
                    float ge = s.irr_wE[(int64_t)q_irr * 7 + q] * r;
                    if (q == 0) {
                        uint8_t nlr = s.irr_nl[q_irr];
                        if (nlr) {
                            float Ec = s.E[c];
                            ge = fmaf(s.irr_nlw[q_irr] * (nlr == 1 ? nl_deriv(s.nonlinear_m, s.nl_coef_m, Ec)
                                                                   : nl_deriv(s.nonlinear_p, s.nl_coef_p, Ec)), r, ge);
                        }
                    }
                    s.gE[c] = ge * s.inv_n_points;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            float gq = w[q] * r;
            if (q == 0 && s.nl)
                gq = fmaf(nlw0 * nl_deriv(s.nonlinear_m, s.nl_coef_m, u[0]) +
                              nlw1 * nl_deriv(s.nonlinear_p, s.nl_coef_p, u[0]), r, gq);
            if (!S4) G7[q * N + p] = gq;
            else if (q == 0) G7[i4[0]] = gq;
            else atomicAdd(G7 + i4[q], gq);
        }
    }
    // loss partial: one row per CTA, gradient entries zero
    __shared__ float red[kThreads / 32];
    float v = loss * s.inv_n_points;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (loss_strip) {   // one float per CTA; the reduction kernel adds the strip to the loss column
        if (threadIdx.x == 0) {
            float t = 0.0f;
            for (int wv = 0; wv < kThreads / 32; ++wv) t += red[wv];
            loss_strip[blockIdx.x] = t;
        }
        return;
    }
    float* row = s.partials + (size_t)(row0 + blockIdx.x) * np1;
    for (int i = threadIdx.x; i < np1 - 1; i += kThreads) row[i] = 0.0f;
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int wv = 0; wv < kThreads / 32; ++wv) t += red[wv];
        row[np1 - 1] = t;
    }
}

// Z2: backward through the extrapolation of the crossed sites of the batch
template <class NET>
__global__ void __launch_bounds__(kThreads, 1) points_extrap_bwd_kernel(PointsArgs a, int row0, int stride) {
    const nbm_points_step_t& s = a.s;
    typename NET::Acc acc;
    acc.zero();
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < s.n_crossed; c += nwarps) {
        int64_t p = s.c_site[c] % a.n_points;
        if (p < s.p0 || p >= s.p1) continue;
        float ge = s.gE[c];
        if (lane < 27 && ge != 0.0f) {
            float X0 = (float)(lane % 3 - 1) * s.dx, X1 = (float)((lane / 3) % 3 - 1) * s.dy,
                  X2 = (float)(lane / 9 - 1) * s.dz;
            bool plus = (s.c_cube_side[c] >> lane) & 1u;
            float gq = s.B[c * 28 + lane] * ge;
            if (gq != 0.0f)
                NET::grad(plus, s.c_pos[3 * c] + X0, s.c_pos[3 * c + 1] + X1, s.c_pos[3 * c + 2] + X2, gq, acc);
        }
    }
    block_reduce_store<NET>(acc, 0.0f, s.partials + (size_t)row0 * stride, stride);
}

// ---- the 27-cube of a crossed site as a 3 x 3 x 3 mini-lattice --------------------------------------------------------
// A work item is (crossed site c, j = jy + 3 jz): the three cube vertices q = 3 j, 3 j + 1, 3 j + 2 share (y, z) and march
// along x (get_Xijk: x fastest, discretization.py:164-197), so the lattice kernels' machinery applies: the x-independent
// part of the first layer is computed once per item, the forward evaluates the three nodes together, the backward is the
// pair-split gradient of node_grad_kernel.  Replaces the warp-per-site kernels above (lane per vertex, full accumulator
// set per thread at 255 registers): measured at 128^3 zoom 1 (66 k crossed sites) 39 -> fwd and 84 -> bwd microseconds.
constexpr int kCubeFwdThreads = 288;    // 32 sites x 9 items: a site's partial sums meet inside one CTA

template <class NET>
__global__ void __launch_bounds__(kCubeFwdThreads) cube_fwd_kernel(PointsArgs a) {
    const nbm_points_step_t& s = a.s;
    __shared__ float part[kCubeFwdThreads];
    constexpr int HP2 = NET::HPW / 2;
    const int64_t n_sites = s.c_live ? s.n_live : s.n_crossed;     // the batch's own sites, or all of them (filtered)
    for (int64_t c0 = (int64_t)blockIdx.x * 32; c0 < n_sites; c0 += (int64_t)gridDim.x * 32) {
        const int sl = threadIdx.x / 9, j = threadIdx.x - sl * 9;
        int64_t c = c0 + sl;
        float v = 0.0f;
        bool live = c < n_sites;
        if (live) {
            if (s.c_live) c = s.c_live[c];
            const int64_t p = s.c_site[c] % a.n_points;
            live = p >= s.p0 && p < s.p1;
        }
        if (live) {
            const int jy = j % 3, jz = j / 3;
            const float px = s.c_pos[3 * c], y = s.c_pos[3 * c + 1] + (float)(jy - 1) * s.dy,
                        z = s.c_pos[3 * c + 2] + (float)(jz - 1) * s.dz;
            const float xs[3] = {px + -1.0f * s.dx, px + 0.0f * s.dx, px + 1.0f * s.dx};
            const unsigned side3 = (s.c_cube_side[c] >> (3 * j)) & 7u;
            // branch-free over the sides: a warp at the interface holds plus AND minus vertices, so every item takes the
            // paired plus-head evaluation of its three vertices (wasted on minus vertices, but one pass instead of the two
            // or three a divergent warp would run) and the tiny minus head where a vertex needs it
            float u[3];
            u64 yz[HP2];
            NET::first_layer_yz(y, z, yz);
            NET::P::template forward_many<0, 3>(xs, yz, u);
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (!((side3 >> i) & 1u)) {
                    float am[NET::LMD][NET::HMW];
                    u[i] = NET::M::template forward<NET::P::NP>(xs[i], y, z, am);
                }
            const float* B = s.B + c * 28 + 3 * j;
            v = fmaf(B[2], u[2], fmaf(B[1], u[1], B[0] * u[0]));
        }
        part[threadIdx.x] = v;
        __syncthreads();
        if (j == 0 && live) {
            float e = part[threadIdx.x];
#pragma unroll
            for (int k = 1; k < 9; ++k) e += part[threadIdx.x + k];
            s.E[c] = e + s.B[c * 28 + 27];
            s.gE[c] = 0.0f;
        }
        __syncthreads();
    }
}

template <class NET>
__global__ void __launch_bounds__(kGradThreads, 1) cube_grad_kernel(PointsArgs a, int row0, int stride) {
    const nbm_points_step_t& s = a.s;
    using P = typename NET::P;
    using M = typename NET::M;
    constexpr int H = NET::HPW, HP2 = H / 2, NP = NET::NP;
    extern __shared__ __align__(16) float dsm[];
    float* hs = dsm + threadIdx.x;                      // [3H] hoisted sums of this thread, stride kGradThreads
    float* red = dsm + 3 * H * kGradThreads;            // [warps][NP + 1]
#pragma unroll
    for (int i = 0; i < 3 * H; ++i) hs[i * kGradThreads] = 0.0f;
    typename P::AccS acc;
    acc.zero();
    float accm[M::NP];
#pragma unroll
    for (int i = 0; i < M::NP; ++i) accm[i] = 0.0f;
    const bool par = (threadIdx.x & 1) != 0;
    // equal contiguous ranges of the (site, j) items per CTA; every thread of the CTA runs the same number of rounds
    const int64_t total = (s.c_live ? s.n_live : s.n_crossed) * 9;
    const int64_t lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
    for (int64_t base = lo; base < hi; base += kGradThreads) {
        const int64_t it = base + threadIdx.x;
        float y = 0.0f, z = 0.0f, px = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
        unsigned side3 = 0u;
        if (it < hi) {
            // every load of the item is issued before any of them is used: one memory round trip per round, not four
            const int64_t ci = it / 9;
            const int j = (int)(it - ci * 9), jy = j % 3, jz = j / 3;
            const int64_t c = s.c_live ? (int64_t)s.c_live[ci] : ci;
            const int64_t site = s.c_site[c];
            const float ge0 = s.gE[c];
            const float cx = s.c_pos[3 * c], cy = s.c_pos[3 * c + 1], cz = s.c_pos[3 * c + 2];
            const unsigned cs = s.c_cube_side[c];
            const float* B = s.B + c * 28 + 3 * j;
            const float b0 = B[0], b1 = B[1], b2 = B[2];
            const int64_t p = site % a.n_points;
            const float ge = (p >= s.p0 && p < s.p1) ? ge0 : 0.0f;    // sites of other batches contribute nothing
            px = cx;
            y = cy + (float)(jy - 1) * s.dy;
            z = cz + (float)(jz - 1) * s.dz;
            side3 = (cs >> (3 * j)) & 7u;
            g0 = b0 * ge; g1 = b1 * ge; g2 = b2 * ge;
        }
        u64 yz[HP2];
        P::template first_layer_yz<0>(y, z, yz);
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {     // (one copy of the backward code: the three nodes run through it in turn)
            const float x = px + (float)(i - 1) * s.dx, g = i == 0 ? g0 : (i == 1 ? g1 : g2);
            const bool plus = (side3 >> i) & 1u;
            const bool do_p = plus && g != 0.0f;
            if (__any_sync(0xffffffffu, do_p)) P::template grad_split<false>(x, yz, do_p ? g : 0.0f, acc, par);
            if (!plus && g != 0.0f) {
                float am[NET::LMD][NET::HMW];
                M::template forward<P::NP>(x, y, z, am);
                M::template backward<P::NP, M::NP, P::NP>(x, y, z, am, g, accm);
            }
        }
        // fold the item's sum(delta1) into the bias and the y, z rows of the first layer
#pragma unroll
        for (int jj = 0; jj < H; ++jj) {
            const float t = (jj & 1) ? hi32(acc.t[jj / 2]) : lo32(acc.t[jj / 2]);
            hs[jj * kGradThreads] += t;
            hs[(H + jj) * kGradThreads] = fmaf(y, t, hs[(H + jj) * kGradThreads]);
            hs[(2 * H + jj) * kGradThreads] = fmaf(z, t, hs[(2 * H + jj) * kGradThreads]);
        }
#pragma unroll
        for (int jj = 0; jj < HP2; ++jj) acc.t[jj] = 0ull;
    }
    // block reduction into one partial row (as node_grad_kernel; the loss column stays 0)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i <= NP; ++i) {
        float val;
        if (i < P::NP) val = P::split_get(acc, i, par, hs, kGradThreads);
        else if (i < NP) val = accm[i < NP ? i - P::NP : 0];
        else val = 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) red[warp * (NP + 1) + i] = val;
    }
    __syncthreads();
    float* row = s.partials + (size_t)(row0 + blockIdx.x) * stride;
    for (int i = threadIdx.x; i < NP + 1; i += kGradThreads) {
        float val = 0.0f;
#pragma unroll
        for (int w = 0; w < kGradThreads / 32; ++w) val += red[w * (NP + 1) + i];
        row[i < NP ? i : stride - 1] = val;
    }
}

template <class NET>
static int launch_points(const nbm_points_step_t& s, cudaStream_t st) {
    const int sms = sm_count();
    const int np1 = NET::NP + 1;
    PointsArgs a;
    a.s = s;
    const float sh[7][3] = {{0, 0, 0}, {-s.dx, 0, 0}, {s.dx, 0, 0}, {0, -s.dy, 0}, {0, s.dy, 0}, {0, 0, -s.dz}, {0, 0, s.dz}};
    for (int q = 0; q < 7; ++q)
        for (int c = 0; c < 3; ++c) a.shift[q][c] = sh[q][c];
    a.n_points = (int64_t)s.nx * s.ny * s.nz;
    const int64_t nb = s.p1 - s.p0;
    const int plane = s.ny * s.nz;
    // the x planes that hold the batch
    NodeView v;
    v.xe = s.xs7; v.ye = s.ys7; v.ze = s.zs7;
    v.ex = s.nx; v.ey = s.ny; v.ez = s.nz;
    v.x_begin = (int)(s.p0 / plane);
    v.x_end = (int)((s.p1 + plane - 1) / plane);
    v.lo = s.p0; v.hi = s.p1;
    v.rep_nodes = a.n_points;
    v.nrep = 7;
    v.side = s.side; v.U = s.U7; v.G = s.G7; v.R = nullptr; v.Gz = nullptr;
    v.inv_n = s.inv_n_points; v.partials = s.partials;
    // zoom level 1 on the 4 shared lattices (padded dims; the batch is whole x planes; one more plane for the x-half lattice)
    const bool s4 = s.U4 != nullptr;
    if (s4) {
        v.xe = s.xs4; v.ye = s.ys4; v.ze = s.zs4;
        v.ex = s.nx + 1; v.ey = s.ny + 1; v.ez = s.nz + 1;
        v.x_end = v.x_begin + (int)(nb / plane) + 1;
        v.rep_nodes = (int64_t)v.ex * v.ey * v.ez;
        v.lo = 0; v.hi = v.rep_nodes;
        v.nrep = 4;
        v.side = s.side4; v.U = s.U4; v.G = s.G4; v.Gz = s.G4;
    }
    const int nrep = v.nrep;
    int xchunk = 16;
    Tasks T = make_tasks(v.x_end - v.x_begin, v.ey, v.ez, xchunk);
    const int64_t itersF = (int64_t)T.mblocks * (v.x_end - v.x_begin) * nrep;
    int perF = 12;      // forward CTAs per SM (3 resident), fewer while a range would hold < 24 plane iterations
    while (perF > 3 && itersF / ((int64_t)sms * perF) < 24) perF -= 3;
    const int gridF = (int)min(itersF, (int64_t)sms * perF);
    Tasks Tg = make_tasks(v.x_end - v.x_begin, v.ey, v.ez, xchunk, kGradThreads);
    // one gradient CTA per SM: the replicas are one range set, every CTA gets the same share (one even wave)
    const int64_t itersG = (int64_t)Tg.mblocks * (v.x_end - v.x_begin) * nrep;
    const int gridG = (int)min(itersG, (int64_t)sms);
    Tg.split = run_split(itersG, gridG);
    // rows kernel: HBM/latency bound (92 B per point, 14 strided loads per thread): 8 CTAs per SM keep enough loads in flight
    static const int rows_per_sm = getenv("NBM_ROWS_CTAS") ? atoi(getenv("NBM_ROWS_CTAS")) : 6;
    const int gridR = (int)min((int64_t)sms * rows_per_sm, (nb + kThreads - 1) / kThreads);
    int gridE = 0;
    const int64_t n_sites = s.c_live ? s.n_live : s.n_crossed;      // crossed sites the cube kernels walk
    if (n_sites > 0) gridE = (int)min((int64_t)sms, (n_sites * 32 + kThreads - 1) / kThreads);
    const bool pc = s.coef26 != nullptr;
    const int n_pc = pc ? PrecondNet<8, 4>::NP : 0;
    const int stride = NET::NP + n_pc + 1;
    const int gridP = pc ? min(s.n_pc_rows, 3 * sms) : 0;
    // partial rows: [gridG gradient rows][gridR loss rows][gridE cube rows][gridP preconditioner rows]; without a
    // preconditioner the rows kernel's per-CTA loss sums are ONE compact strip of gridR floats behind the cube rows
    // instead of gridR mostly-zero rows (the reduction then reads 296 rows instead of up to 1184)
    const bool strip = !pc;
    const int rowsR = strip ? 0 : gridR, strip_rows = strip ? (gridR + stride - 1) / stride : 0;
    const int rows_needed = gridG + rowsR + gridE + gridP + strip_rows;
    float* loss_strip = strip ? s.partials + (size_t)(gridG + gridE) * stride : nullptr;
    if (rows_needed > s.n_partial_rows) {
        set_error("partials buffer has %d rows, %d needed", s.n_partial_rows, rows_needed);
        return NBM_ERR_WORKSPACE;
    }
    // plans of one level share the partials buffer and their row roles differ with the batch size: with the wider
    // preconditioner rows every kernel writes only its own columns, so start from zero
    if (pc) cudaMemsetAsync(s.partials, 0, sizeof(float) * (size_t)rows_needed * stride, st);
    static const bool cube_kernels = !(getenv("NBM_CUBE_KERNELS") && getenv("NBM_CUBE_KERNELS")[0] == '0');
    if (n_sites > 0 && cube_kernels)
        cube_fwd_kernel<NET><<<(unsigned)min((int64_t)sms * 8, (n_sites + 31) / 32), kCubeFwdThreads, 0, st>>>(a);
    else if (s.n_crossed > 0 && !cube_kernels)
        points_extrap_kernel<NET><<<(unsigned)min((int64_t)sms * 4, (s.n_crossed * 32 + kThreads - 1) / kThreads),
                                    kThreads, 0, st>>>(a);
    fwd_nodes_kernel<NET, true><<<gridF, kThreads, 0, st>>>(v, T);
    if (pc)
        precond_fwd_kernel<8, 4><<<(unsigned)min((int64_t)sms * 8, (nb + kThreads - 1) / kThreads), kThreads, 0, st>>>(
            s.coef26 + s.p0, a.n_points, nb, s.pc_params, s.pc_scale, s.Pc + s.p0);
    if (s4) points_rows_kernel<true><<<gridR, kThreads, 0, st>>>(a, s.U4, s.G4, gridG, stride, loss_strip);
    else points_rows_kernel<false><<<gridR, kThreads, 0, st>>>(a, s.U7, s.G7, gridG, stride, loss_strip);
    if (pc)   // loss + d loss/d theta_P from the raw residuals kept in `rows`
        precond_kernel<8, 4><<<gridP, kThreads, 0, st>>>(s.coef26 + s.p0, a.n_points, s.rows + s.p0, nullptr, nb, s.pc_params,
                                                         s.pc_scale, s.inv_n_points,
                                                         s.partials + (size_t)(gridG + rowsR + gridE) * stride, stride,
                                                         NET::NP, NET::NP + n_pc);
    v.row0 = 0; v.row_stride = pc ? stride : 0; v.loss_col = NET::NP + n_pc;
    {
        cudaError_t e = launch_node_grad<NET, true>(dim3(gridG), v, Tg, st);
        if (e != cudaSuccess) return cuda_check(e, "node_grad attribute");
    }
    if (gridE > 0 && cube_kernels) {
        static unsigned long long configured = 0ull;   // per instantiation and device
        int dev = 0;
        cudaGetDevice(&dev);
        constexpr int bytes = grad_smem_bytes<NET>();
        if (!((configured >> (dev & 63)) & 1ull)) {
            cudaError_t e = cudaFuncSetAttribute(cube_grad_kernel<NET>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            if (e != cudaSuccess) return cuda_check(e, "cube_grad attribute");
            configured |= 1ull << (dev & 63);
        }
        cube_grad_kernel<NET><<<gridE, kGradThreads, bytes, st>>>(a, gridG + rowsR, stride);
    } else if (gridE > 0) {
        points_extrap_bwd_kernel<NET><<<gridE, kThreads, 0, st>>>(a, gridG + rowsR, stride);
    }
    reduce_partials_kernel<<<(stride * 32 + 127) / 128, 128, 0, st>>>(s.partials, gridG + rowsR + gridE + gridP, stride, s.loss_grad,
                                                                      loss_strip, strip ? gridR : 0);
    return cuda_check(cudaGetLastError(), "points step launch");
}

static int dispatch_points(const nbm_points_step_t& s, cudaStream_t st) { NBM_NET_DISPATCH(s.net, launch_points, s, st); }

template <int LP, int HP, int LM, int HM>
static int launch_eval_impl(const nbm_lvl_t& L, const float* pts, int64_t n, float dx, float dy, float dz, float* u,
                            float* grad_u, float* grad_n, cudaStream_t st) {
    evaluate_kernel<Net<LP, HP, LM, HM>, LP, HP, LM, HM>
        <<<(unsigned)((n + 127) / 128), 128, 0, st>>>(L, pts, n, dx, dy, dz, u, grad_u, grad_n);
    return cuda_check(cudaGetLastError(), "evaluate launch");
}

static int net_num_params(const nbm_net_t& n) {
    auto cnt = [](int L, int H) { return 3 * H + H + (L - 1) * (H * H + H) + H + 1; };
    return cnt(n.layers_p, n.hidden_p) + cnt(n.layers_m, n.hidden_m);
}

}  // namespace nbm

using namespace nbm;

extern "C" {

int nbm_net_num_params(const nbm_net_t* net) {
    if (!net || net->layers_p < 1 || net->layers_m < 1 || net->hidden_p < 1 || net->hidden_m < 1) return -1;
    return net_num_params(*net);
}

int nbm_upload_params(const nbm_net_t* net, const float* params, nbm_stream_t stream) {
    NBM_REQUIRE(net && params, "null pointer");
    int P = nbm_net_num_params(net);
    if (P <= 0 || P > NBM_MAXP) {
        set_error("parameter count %d outside (0, %d]", P, NBM_MAXP);
        return NBM_ERR_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    float* stage = nullptr;
    int rc = cuda_check(cudaGetSymbolAddress((void**)&stage, g_stage), "staging buffer");
    if (rc) return rc;
    prep_params_kernel<<<(P + 127) / 128, 128, 0, st>>>(*net, params, stage, P);
    NBM_LAUNCH_CHECK("prep_params");
    return cuda_check(cudaMemcpyToSymbolAsync(c_P, stage, sizeof(float) * 3 * NBM_MAXP, 0, cudaMemcpyDeviceToDevice, st),
                      "upload params");
}

int nbm_step_partial_rows(void) { return kPartialRows; }
int nbm_precond_num_params(int d1, int d2) { return 26 * d1 + d1 + d1 * d2 + d2 + d2 + 1; }

int nbm_ffma_probe_f32(int iters, float* out, double* flops_host, nbm_stream_t stream) {
    NBM_REQUIRE(out && iters > 0, "bad arguments");
    int blocks = sm_count() * 8;
    ffma_probe_kernel<<<blocks, 256, 0, as_stream(stream)>>>(iters, out);
    if (flops_host) *flops_host = 2.0 * 16.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    NBM_LAUNCH_CHECK("ffma probe");
    return NBM_OK;
}

int nbm_loss_grad_shared_f32(const nbm_shared_step_t* s, nbm_stream_t stream) {
    NBM_REQUIRE(s, "null plan");
    NBM_REQUIRE(s->xe && s->ye && s->ze && s->side && s->rhs, "null tables");
    if (s->faces) {
        NBM_REQUIRE(s->cface && s->dinv, "null face tables");
        NBM_REQUIRE(s->n_irr == 0 || (s->irr_wU && s->irr_rhs), "null irregular-row tables (faces mode)");
        NBM_REQUIRE(((s->ey * s->ez) % 4 == 0) && (s->ez % 2 == 0) &&
                        ((((uintptr_t)s->cface | (uintptr_t)s->dinv | (uintptr_t)s->rhs | (uintptr_t)s->U |
                           (uintptr_t)s->R | (uintptr_t)s->G | (uintptr_t)s->nl | (uintptr_t)s->kv) & 15) == 0),
                    "faces mode needs plane % 4 == 0, even ez and 16-byte aligned arrays");
    } else {
        NBM_REQUIRE(s->w, "null tables");
    }
    if (s->g_ptr) {
        NBM_REQUIRE(s->g_ent && s->list_nodes && s->n_list >= 0, "null list-node tables");
        NBM_REQUIRE(s->n_crossed == 0 || (s->ge_ptr && s->ge_ent), "null crossed-site incidence");
        NBM_REQUIRE(!s->S, "the gathered list adjoint and the fused adjoint (S) are alternatives");
    }
    NBM_REQUIRE(!s->irr_c_soa == !s->irr_wE_soa && !s->irr_c_soa == !s->irr_wU_soa, "transposed irregular-row tables come as a set");
    if (s->G2 || s->Rq) {
        NBM_REQUIRE(s->G2 && s->Rq && s->list_nodes && s->n_list >= 0 && !s->g_ptr,
                    "the list chain beside the stencil needs G2, Rq and list_nodes (and no gathered adjoint)");
    }
    if (s->coef26) {
        NBM_REQUIRE(s->pc_params, "null preconditioner parameters");
        NBM_REQUIRE(s->n_pc_rows >= 1 && s->n_partial_rows > s->n_pc_rows, "no partial rows for the preconditioner");
        NBM_REQUIRE(!(s->pc_nodes_m || s->pc_nodes_p) || s->n_pc_rows >= 3, "the per-side preconditioner kernels need n_pc_rows >= 3");
        NBM_REQUIRE((s->n_pc_m == 0 || s->pc_nodes_m) && (s->n_pc_p == 0 || s->pc_nodes_p), "null preconditioner node lists");
        NBM_REQUIRE(!s->S, "the fused adjoint path does not take a preconditioner");
        if (s->pc_d1 != 8 || s->pc_d2 != 4) {
            set_error("preconditioner widths (%d, %d) are outside the compiled kernel set ((8, 4))", s->pc_d1, s->pc_d2);
            return NBM_ERR_UNSUPPORTED;
        }
    }
    NBM_REQUIRE(s->ex >= 3 && s->ey >= 3 && s->ez >= 3, "lattice too small");
    NBM_REQUIRE(s->U && s->R && s->G && s->partials && s->loss_grad, "null work buffers");
    NBM_REQUIRE(s->n_partial_rows >= 1, "n_partial_rows must be >= 1");
    NBM_REQUIRE(s->n_crossed == 0 || (s->c_node && s->B && s->E && s->gE), "null crossed-site tables");
    NBM_REQUIRE(s->n_irr == 0 || (s->irr_point && s->irr_wE && s->irr_c && s->irr_nl && s->irr_nlw),
                "null irregular-row tables");
    NBM_REQUIRE((s->nonlinear_m == NBM_NL_NONE && s->nonlinear_p == NBM_NL_NONE) || s->nl,
                "nonlinear operator needs the nl table");
    return dispatch_shared(*s, as_stream(stream));
}

int nbm_loss_grad_points_f32(const nbm_points_step_t* s, nbm_stream_t stream) {
    NBM_REQUIRE(s, "null plan");
    NBM_REQUIRE(s->xs && s->ys && s->zs && s->side && s->w && s->rhs && s->irr, "null tables");
    NBM_REQUIRE(s->nx > 0 && s->ny > 0 && s->nz > 0, "empty grid");
    NBM_REQUIRE(s->p0 >= 0 && s->p1 > s->p0 && s->p1 <= (int64_t)s->nx * s->ny * s->nz, "bad batch range");
    NBM_REQUIRE(s->dx > 0 && s->dy > 0 && s->dz > 0, "cell size must be positive");
    NBM_REQUIRE(s->partials && s->loss_grad && s->n_partial_rows >= 2, "null work buffers");
    NBM_REQUIRE(s->U7 && s->G7 && s->xs7 && s->ys7 && s->zs7, "null site work buffers / displaced coordinate arrays");
    if (s->U4 || s->G4 || s->xs4 || s->ys4 || s->zs4 || s->side4) {
        NBM_REQUIRE(s->U4 && s->G4 && s->xs4 && s->ys4 && s->zs4 && s->side4, "the 4 shared lattices need all of xs4, ys4, zs4, side4, U4, G4");
        const int64_t plane = (int64_t)s->ny * s->nz;
        NBM_REQUIRE(s->p0 % plane == 0 && s->p1 % plane == 0, "the 4 shared lattices serve batches of whole x planes");
    }
    NBM_REQUIRE(s->n_crossed == 0 || (s->c_site && s->c_pos && s->c_cube_side && s->B && s->E && s->gE),
                "null crossed-site tables");
    NBM_REQUIRE(s->n_irr == 0 || (s->irr_wE && s->irr_c && s->irr_nl && s->irr_nlw), "null irregular-row tables");
    NBM_REQUIRE((s->nonlinear_m == NBM_NL_NONE && s->nonlinear_p == NBM_NL_NONE) || s->nl,
                "nonlinear operator needs the nl table");
    if (s->coef26) {
        NBM_REQUIRE(s->pc_params && s->Pc && s->rows, "the preconditioner needs pc_params, Pc and rows");
        NBM_REQUIRE(s->n_pc_rows >= 1, "no partial rows for the preconditioner");
        if (s->pc_d1 != 8 || s->pc_d2 != 4) {
            set_error("preconditioner widths (%d, %d) are outside the compiled kernel set ((8, 4))", s->pc_d1, s->pc_d2);
            return NBM_ERR_UNSUPPORTED;
        }
    }
    return dispatch_points(*s, as_stream(stream));
}


static unsigned long long g_comm_timeout_ns = 30ull * 1000000000ull;

int nbm_comm_set_timeout(double seconds) {
    NBM_REQUIRE(seconds >= 0.0 && seconds < 1e9, "timeout must be >= 0 seconds (0 = wait without a bound)");
    g_comm_timeout_ns = (unsigned long long)(seconds * 1e9);
    return NBM_OK;
}

int nbm_comm_alloc_local(void** local_block) {
    NBM_REQUIRE(local_block, "null pointer");
    void* p = nullptr;
    int rc = cuda_check(cudaMalloc(&p, sizeof(CommBlock)), "cudaMalloc(comm block)");
    if (rc) return rc;
    rc = cuda_check(cudaMemset(p, 0, sizeof(CommBlock)), "memset(comm block)");
    if (rc) return rc;
    *local_block = p;
    return NBM_OK;
}

int nbm_enable_peer_access(int device, int peer) {
    NBM_REQUIRE(device >= 0 && peer >= 0 && device != peer, "bad device pair");
    int can = 0;
    int rc = cuda_check(cudaDeviceCanAccessPeer(&can, device, peer), "cudaDeviceCanAccessPeer");
    if (rc) return rc;
    if (!can) {
        set_error("device %d cannot access device %d's memory (no NVLink / PCIe peer path)", device, peer);
        return NBM_ERR_UNSUPPORTED;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    cudaSetDevice(cur);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return NBM_OK;
    }
    return cuda_check(e, "cudaDeviceEnablePeerAccess");
}

int nbm_comm_alloc(void** local_block, unsigned char handle[NBM_IPC_HANDLE_BYTES]) {
    NBM_REQUIRE(local_block && handle, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == NBM_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    int rc = cuda_check(cudaMalloc(&p, sizeof(CommBlock)), "cudaMalloc(comm block)");
    if (rc) return rc;
    rc = cuda_check(cudaMemset(p, 0, sizeof(CommBlock)), "memset(comm block)");
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    rc = cuda_check(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
    if (rc) return rc;
    memcpy(handle, &h, sizeof(h));
    *local_block = p;
    return NBM_OK;
}

int nbm_comm_open_peer(const unsigned char handle[NBM_IPC_HANDLE_BYTES], void** peer_block) {
    NBM_REQUIRE(handle && peer_block, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return cuda_check(cudaIpcOpenMemHandle(peer_block, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

int nbm_comm_close_peer(void* peer_block) { return cuda_check(cudaIpcCloseMemHandle(peer_block), "cudaIpcCloseMemHandle"); }
int nbm_comm_free(void* local_block) { return cuda_check(cudaFree(local_block), "cudaFree(comm block)"); }

int nbm_comm_error(void* local_block) {
    if (!local_block) return -1;
    unsigned int e = 0;
    if (cudaMemcpy(&e, &reinterpret_cast<CommBlock*>(local_block)->error, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    return (int)e;
}

int nbm_reduce_allreduce_f32(const float* partials, int rows, int np1, int rank, int world, void* const* blocks_host,
                             int32_t* step_dev, float* out, nbm_stream_t stream) {
    NBM_REQUIRE(partials && blocks_host && step_dev && out, "null pointer");
    NBM_REQUIRE(rows > 0 && np1 > 0 && np1 <= NBM_MAXP + 1, "bad sizes");
    NBM_REQUIRE(world >= 1 && world <= NBM_COMM_MAX_RANKS && rank >= 0 && rank < world, "bad rank/world");
    CommPeers peers;
    for (int r = 0; r < NBM_COMM_MAX_RANKS; ++r) peers.b[r] = r < world ? reinterpret_cast<CommBlock*>(blocks_host[r]) : nullptr;
    for (int r = 0; r < world; ++r) NBM_REQUIRE(peers.b[r], "null peer block");
    const int threads = 1024;   // np1 <= 1025 columns... one column per thread in the last phase needs np1 <= 1024
    NBM_REQUIRE(np1 <= 1024, "np1 must be <= 1024");
    const int ngrp = threads / np1 > 0 ? threads / np1 : 1;
    cudaError_t le = launch_pdl(reduce_allreduce_kernel, dim3(1), dim3(threads), sizeof(float) * (size_t)ngrp * np1, as_stream(stream),
                                partials, rows, np1, rank, world, peers, step_dev, out, g_comm_timeout_ns);
    if (le != cudaSuccess) return cuda_check(le, "reduce_allreduce launch");
    NBM_LAUNCH_CHECK("reduce_allreduce");
    return NBM_OK;
}

int nbm_reduce_allreduce_finalize_f32(const nbm_optimizer_t* opt, const nbm_net_t* net, const float* partials, int rows,
                                      int np1, int rank, int world, void* const* blocks_host, int32_t* step_dev,
                                      float* loss_grad, float* params, float* state, int32_t* count, float* loss_hist,
                                      nbm_stream_t stream) {
    NBM_REQUIRE(opt && net && partials && blocks_host && step_dev && loss_grad && params && state && count, "null pointer");
    NBM_REQUIRE(opt->n_params > 0 && opt->n_params <= NBM_MAXP, "bad parameter count");
    if (opt->optimizer < 0 || opt->optimizer > 2 || opt->scheduler < 0 || opt->scheduler > 1) {
        set_error("unknown optimizer id %d", opt->optimizer);
        return NBM_ERR_UNSUPPORTED;
    }
    const int n_net = nbm_net_num_params(net);
    NBM_REQUIRE(n_net > 0 && n_net <= opt->n_params, "the optimizer must cover at least the network's parameters");
    NBM_REQUIRE(rows > 0 && np1 == opt->n_params + 1 && np1 <= 1024, "partial rows must be n_params + 1 <= 1024 floats wide");
    NBM_REQUIRE(world >= 1 && world <= NBM_COMM_MAX_RANKS && rank >= 0 && rank < world, "bad rank/world");
    CommPeers peers;
    for (int r = 0; r < NBM_COMM_MAX_RANKS; ++r) peers.b[r] = r < world ? reinterpret_cast<CommBlock*>(blocks_host[r]) : nullptr;
    for (int r = 0; r < world; ++r) NBM_REQUIRE(peers.b[r], "null peer block");
    float* stage = nullptr;
    int rc = cuda_check(cudaGetSymbolAddress((void**)&stage, g_stage), "staging buffer");
    if (rc) return rc;
    const int threads = 1024;
    const int ngrp = threads / np1 > 0 ? threads / np1 : 1;
    cudaError_t le = launch_pdl(reduce_allreduce_finalize_kernel, dim3(1), dim3(threads), sizeof(float) * (size_t)ngrp * np1,
                                as_stream(stream), partials, rows, np1, rank, world, peers, step_dev, loss_grad, g_comm_timeout_ns,
                                *opt, *net, n_net, params, state, count, loss_hist, stage);
    if (le != cudaSuccess) return cuda_check(le, "reduce_allreduce_finalize launch");
    NBM_LAUNCH_CHECK("reduce_allreduce_finalize");
    return NBM_OK;
}

int nbm_apply_update_f32(const nbm_optimizer_t* opt, const float* loss_grad, float* params, float* state,
                         int32_t* count, float* loss_hist, nbm_stream_t stream) {
    NBM_REQUIRE(opt && loss_grad && params && state && count, "null pointer");
    NBM_REQUIRE(opt->n_params > 0 && opt->n_params <= NBM_MAXP, "bad parameter count");
    if (opt->optimizer < 0 || opt->optimizer > 2 || opt->scheduler < 0 || opt->scheduler > 1) {
        set_error("unknown optimizer id %d", opt->optimizer);
        return NBM_ERR_UNSUPPORTED;
    }
    apply_update_kernel<<<1, 256, 0, as_stream(stream)>>>(*opt, loss_grad, params, state, count, loss_hist);
    NBM_LAUNCH_CHECK("apply_update");
    return NBM_OK;
}

int nbm_finalize_step_f32(const nbm_optimizer_t* opt, const nbm_net_t* net, const float* partials, int rows,
                          int row_stride, float* loss_grad, float* params, float* state, int32_t* count,
                          float* loss_hist, nbm_stream_t stream) {
    NBM_REQUIRE(opt && net && loss_grad && params && state && count, "null pointer");
    NBM_REQUIRE(opt->n_params > 0 && opt->n_params <= NBM_MAXP, "bad parameter count");
    if (opt->optimizer < 0 || opt->optimizer > 2 || opt->scheduler < 0 || opt->scheduler > 1) {
        set_error("unknown optimizer id %d", opt->optimizer);
        return NBM_ERR_UNSUPPORTED;
    }
    const int n_net = nbm_net_num_params(net);
    NBM_REQUIRE(n_net > 0 && n_net <= opt->n_params, "the optimizer must cover at least the network's parameters");
    NBM_REQUIRE(!partials || (rows > 0 && row_stride == opt->n_params + 1 && row_stride <= 1024),
                "partial rows must be n_params + 1 <= 1024 floats wide");
    cudaStream_t st = as_stream(stream);
    float* stage = nullptr;
    int rc = cuda_check(cudaGetSymbolAddress((void**)&stage, g_stage), "staging buffer");
    if (rc) return rc;
    const int threads = 1024;
    const int ngrp = partials ? (threads / row_stride > 0 ? threads / row_stride : 1) : 0;
    cudaError_t le = launch_pdl(finalize_step_kernel, dim3(1), dim3(threads),
                                sizeof(float) * (size_t)ngrp * (partials ? row_stride : 0), st, *opt, *net, n_net, partials, rows,
                                row_stride, loss_grad, params, state, count, loss_hist, stage);
    if (le != cudaSuccess) return cuda_check(le, "finalize_step launch");
    NBM_LAUNCH_CHECK("finalize_step");
    return NBM_OK;
}

int nbm_upload_staged_params(nbm_stream_t stream) {
    float* stage = nullptr;
    int rc = cuda_check(cudaGetSymbolAddress((void**)&stage, g_stage), "staging buffer");
    if (rc) return rc;
    return cuda_check(cudaMemcpyToSymbolAsync(c_P, stage, sizeof(float) * 3 * NBM_MAXP, 0, cudaMemcpyDeviceToDevice,
                                              as_stream(stream)),
                      "upload staged params");
}

int nbm_evaluate_f32(const nbm_net_t* net, const nbm_lvl_t* lvl, const float* pts, int64_t n, float dx, float dy,
                     float dz, float* u, float* grad_u, float* grad_n, nbm_stream_t stream) {
    NBM_REQUIRE(net && lvl && (lvl->eval_phi || (lvl->phi_g && lvl->xg && lvl->yg && lvl->zg)), "null pointer");
    NBM_REQUIRE(n >= 0, "negative n");
    if (n == 0) return NBM_OK;
    NBM_REQUIRE(pts && u, "null pointer");
    cudaStream_t st = as_stream(stream);
    const nbm_net_t& q = *net;
    if (q.layers_p == 2 && q.hidden_p == 10 && q.layers_m == 1 && q.hidden_m == 1)
        return launch_eval_impl<2, 10, 1, 1>(*lvl, pts, n, dx, dy, dz, u, grad_u, grad_n, st);
    if (q.layers_p == 2 && q.hidden_p == 10 && q.layers_m == 1 && q.hidden_m == 3)
        return launch_eval_impl<2, 10, 1, 3>(*lvl, pts, n, dx, dy, dz, u, grad_u, grad_n, st);
    if (q.layers_p == 1 && q.hidden_p == 10 && q.layers_m == 1 && q.hidden_m == 1)
        return launch_eval_impl<1, 10, 1, 1>(*lvl, pts, n, dx, dy, dz, u, grad_u, grad_n, st);
    if (q.layers_p == 3 && q.hidden_p == 10 && q.layers_m == 1 && q.hidden_m == 1)
        return launch_eval_impl<3, 10, 1, 1>(*lvl, pts, n, dx, dy, dz, u, grad_u, grad_n, st);
    set_error("network shape p(%d x %d) m(%d x %d) is outside the compiled kernel set", q.layers_p, q.hidden_p,
              q.layers_m, q.hidden_m);
    return NBM_ERR_UNSUPPORTED;
}

}  // extern "C"
