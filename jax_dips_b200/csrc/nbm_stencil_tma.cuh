// Residual rows + adjoint stencil of the faces table as ONE kernel (included by nbm_step.cu).
//
//   R[e] = dinv_e ( (sum_f c_f + kv_e) U_e - sum_f c_f U_nb(f) ) - rhs_e     (discretization.py:366-386, regular rows)
//   T    = dinv R on regular rows (0 elsewhere)
//   G[e] = (sum_f c_f + kv_e) T_e - sum_f c_f T_nb(f)  (+ R_e on Dirichlet rows)   = d loss / d U_e, dense part
//
// As two kernels this is 28 + 24 bytes per lattice node of HBM traffic (R written, then R, dinv and the three face
// arrays read again).  Here a CTA owns a (gy rows) x (4 gzq cells) tile of the (y, z) plane and marches along x;
// every x plane of U, the three face-coefficient arrays, dinv, rhs (kv, nl) arrives as a 3-D TMA box
// (cp.async.bulk.tensor.3d, halo rows/columns included, out-of-lattice elements zero-filled by the TMA unit) in a
// shared-memory ring a few planes ahead of the compute, issued by one producer warp; T lives in a 4-plane shared ring
// and never reaches HBM: 28 B/node read, 8 B/node written, one CTA barrier per plane.
//   iteration p:  T[p] on the tile + a 1-cell halo from U[p-1], U[p], U[p+1] and the tables of plane p
//                 (the -x face coefficient is last iteration's +x one, carried in registers), R[p] -> HBM;
//                 barrier; G[p-1] on the tile from T[p-2], T[p-1], T[p] and the coefficients of plane p-1, which the
//                 same thread kept in registers from the previous iteration.
// Arithmetic order is that of residual_faces4_body / adjoint_faces4_body: results are bitwise those of the two kernels.
#include <cuda.h>

namespace stencil_tma {

// CTA = n_main threads (one float4 group of the tile each) + n_halo threads (the T halo: 2 gzq groups of rows -1, gy
// and 2 gy single cells of columns -1, 4 gzq) + the producer warp.  Two sizes: 256 + 96 (one CTA per SM) and
// 128 + 64 (two CTAs per SM: twice as many independent warps to hide the shared-memory and barrier latency).
constexpr int kMainMax = 256, kHaloMax = 96;
constexpr int kThreadsS = kMainMax + kHaloMax + 32;
constexpr int kTRing = 2;

struct Geom {
    int ex, ey, ez;
    int gy, gzq;          // tile: gy rows x gzq groups of 4 cells
    int ny_t, nz_t, nxc;  // tiles in y, z; chunks in x
    int xchunk;           // planes of G per chunk
    int bw, bh;           // box: bh = gy + 4 rows of bw = 4 gzq + 8 floats
    int slot_floats;      // bh * bw rounded up to 128 bytes
    int nsu, nst;         // ring depths: U planes, table planes
    int n_main, n_halo;   // thread roles (multiples of 32)
    int dbg;              // timing experiments: 1 = consumers skip the arithmetic (fill rate of the TMA pipeline alone)
};

struct Maps {
    CUtensorMap U, cx, cy, cz, dinv, rhs, kv, nla, nlb;
};

template <bool KV, bool NL>
__host__ __device__ constexpr int n_tables() { return 5 + (KV ? 1 : 0) + (NL ? 2 : 0); }

template <bool KV, bool NL>
__host__ __device__ inline size_t smem_bytes(const Geom& g) {
    return 128 /* barriers */ + sizeof(float) * ((size_t)g.slot_floats * (g.nsu + (size_t)g.nst * n_tables<KV, NL>()) +
                                                  (size_t)kTRing * (g.gy + 2) * g.bw) + 128 /* alignment slack */;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bar_consumers(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// one row value; same expression order as NBM_ROW of residual_faces4_body
__device__ __forceinline__ float row_val(float di, float rh, float dsum, float u0, float cxm, float uxm, float cxp, float uxp,
                                         float cym, float uym, float cyp, float uyp, float czm, float uzm, float czp, float uzp) {
    float acc = dsum * u0;
    acc = fmaf(-cxm, uxm, acc); acc = fmaf(-cxp, uxp, acc);
    acc = fmaf(-cym, uym, acc); acc = fmaf(-cyp, uyp, acc);
    acc = fmaf(-czm, uzm, acc); acc = fmaf(-czp, uzp, acc);
    return di > 0.f ? fmaf(di, acc, -rh) : (di < 0.f ? u0 - rh : 0.f);
}

// explicit shared-window accesses (32-bit addresses): the compiler cannot prove that pointers derived from the aligned
// dynamic-shared base are shared, and generic LD/ST cost address arithmetic and long-scoreboard stalls
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void mbar_init_a(uint32_t a, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}

template <bool KV, bool NL>
__global__ void __launch_bounds__(kThreadsS, 1)
stencil_tma_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Geom g, nbm_shared_step_t s) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NT = n_tables<KV, NL>();
    // table order inside a table-ring slot
    constexpr int T_CX = 0, T_CY = 1, T_CZ = 2, T_DI = 3, T_RH = 4, T_KV = 5, T_NA = 5 + (KV ? 1 : 0), T_NB = T_NA + 1;
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t nsu = (uint32_t)g.nsu, nst = (uint32_t)g.nst;
    const uint32_t slot_b = (uint32_t)g.slot_floats * 4u;
    // barriers: fullU[nsu] emptyU[nsu] fullT[nst] emptyT[nst]  (<= 16)
    const uint32_t fullU = base, emptyU = base + 8u * nsu, fullT = base + 16u * nsu, emptyT = fullT + 8u * nst;
    const uint32_t ringU = base + 128u, ringT = ringU + nsu * slot_b, Ts = ringT + nst * NT * slot_b;

    const int tid = threadIdx.x;
    const int kMain = g.n_main, kConsumers = g.n_main + g.n_halo;
    const int nt = g.ny_t * g.nz_t;
    const int tile = blockIdx.x % nt, chunk = blockIdx.x / nt;
    const int ty_i = tile / g.nz_t, tz_i = tile - ty_i * g.nz_t;
    const int y0 = ty_i * g.gy, z0 = tz_i * g.gzq * 4;
    const int xa = chunk * g.xchunk, xb = min(g.ex, xa + g.xchunk);
    const int n_iter = xb - xa + 2;           // p = xa-1 .. xb
    const int bw = g.bw;
    const uint32_t bw4 = (uint32_t)bw * 4u;
    const uint32_t box_bytes = (uint32_t)(g.bh * bw * sizeof(float));

    if (tid == 0) {
        for (uint32_t i = 0; i < nsu; ++i) {
            mbar_init_a(fullU + 8u * i, 1);
            mbar_init_a(emptyU + 8u * i, kConsumers / 32);
        }
        for (uint32_t i = 0; i < nst; ++i) {
            mbar_init_a(fullT + 8u * i, 1);
            mbar_init_a(emptyT + 8u * i, kConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    pdl_wait();      // (programmatic launch: U of the forward kernel is complete from here on)
    __syncthreads();

    if (tid >= kConsumers) {
        // ---------------- producer warp: one lane issues the TMA boxes ----------------
        if (tid == kConsumers) {
            const int cz0 = z0 - 4, cy0 = y0 - 2;
            uint32_t iu = 0, ku = 0, it = 0, kt = 0;      // ring slot, wrap parity
            bool wrapped_u = false, wrapped_t = false;
            auto load_u = [&](int px) {
                if (wrapped_u) mbar_wait_a(emptyU + 8u * iu, ku ^ 1u);   // every consumer warp released the previous use
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fullU + 8u * iu), "r"(box_bytes) : "memory");
                tma_load_3d(ringU + iu * slot_b, &maps.U, cz0, cy0, px, fullU + 8u * iu);
                if (++iu == nsu) { iu = 0; ku ^= 1u; wrapped_u = true; }
            };
            auto load_t = [&](int px) {
                if (wrapped_t) mbar_wait_a(emptyT + 8u * it, kt ^ 1u);
                const uint32_t bar = fullT + 8u * it;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes * NT) : "memory");
                const uint32_t d = ringT + it * NT * slot_b;
                tma_load_3d(d + T_CX * slot_b, &maps.cx, cz0, cy0, px, bar);
                tma_load_3d(d + T_CY * slot_b, &maps.cy, cz0, cy0, px, bar);
                tma_load_3d(d + T_CZ * slot_b, &maps.cz, cz0, cy0, px, bar);
                tma_load_3d(d + T_DI * slot_b, &maps.dinv, cz0, cy0, px, bar);
                tma_load_3d(d + T_RH * slot_b, &maps.rhs, cz0, cy0, px, bar);
                if (KV) tma_load_3d(d + T_KV * slot_b, &maps.kv, cz0, cy0, px, bar);
                if (NL) {
                    tma_load_3d(d + T_NA * slot_b, &maps.nla, cz0, cy0, px, bar);
                    tma_load_3d(d + T_NB * slot_b, &maps.nlb, cz0, cy0, px, bar);
                }
                if (++it == nst) { it = 0; kt ^= 1u; wrapped_t = true; }
            };
            load_u(xa - 2);
            load_u(xa - 1);
            for (int n = 0; n < n_iter; ++n) {
                load_u(xa + n);          // U[p+1]
                load_t(xa - 1 + n);      // tables of plane p
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    // role: main thread = group (r, gq) of the tile; halo thread = a group of row -1 / gy, or one cell of column -1 / 4 gzq
    const bool is_main = tid < kMain;
    const int lane = tid & 31;
    int r, gq, cz;        // tile row (-1 .. gy), group in the row, box column of the (first) cell
    bool active, vec;
    if (is_main) {
        r = tid / g.gzq;
        gq = tid - r * g.gzq;
        cz = 4 + 4 * gq;
        active = tid < g.gy * g.gzq;
        vec = true;
    } else {
        const int h = tid - kMain;
        if (h < 2 * g.gzq) {
            r = h < g.gzq ? -1 : g.gy;
            gq = h < g.gzq ? h : h - g.gzq;
            cz = 4 + 4 * gq;
            active = true;
            vec = true;
        } else {
            const int h2 = h - 2 * g.gzq;
            r = h2 < g.gy ? h2 : h2 - g.gy;
            gq = 0;
            cz = h2 < g.gy ? 3 : 4 + 4 * g.gzq;
            active = h2 < 2 * g.gy;
            vec = false;
        }
    }
    if (!active) {   // idle threads of a small tile: harmless addresses, no stores
        r = 0;
        gq = 0;
        cz = 4;
    }
    // z neighbours inside a row come from the adjacent lane (shuffle); the ends of a row / of the warp read shared memory
    const unsigned vmask = __ballot_sync(0xffffffffu, vec);
    const bool left_lane = vec && gq > 0 && lane > 0, right_lane = vec && gq < g.gzq - 1 && lane < 31;
    const int yy = y0 + r, zz = z0 + cz - 4;      // lattice coordinates of the (first) cell
    const bool in_lat = active && yy >= 0 && yy < g.ey && zz >= 0 && zz < g.ez;
    const bool stores = is_main && in_lat;
    const int64_t plane = (int64_t)g.ey * g.ez;
    const int64_t e_yz = (int64_t)yy * g.ez + zz;
    const uint32_t so4 = (uint32_t)((r + 2) * bw + cz) * 4u;   // byte offset of the (first) cell inside a box
    const uint32_t to4 = (uint32_t)((r + 1) * bw + cz) * 4u;   // ... inside a T plane
    const uint32_t tplane_b = (uint32_t)(g.gy + 2) * bw4;
    // iterations whose residual plane this CTA writes: xa <= p < xb and 1 <= p <= ex-2
    const int n_r_lo = max(xa, 1) - (xa - 1), n_r_hi = min(xb, g.ex - 1) - (xa - 1);
    float* pR = s.R + (int64_t)(xa - 1) * plane + e_yz;       // plane p of this thread's cells (running)
    float* pG = s.G + (int64_t)(xa - 2) * plane + e_yz;       // plane p - 1

    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // carried per-cell state: coefficients of plane p-1 (for G[p-1]), own U of planes p-1, p, own T of planes p-2, p-1
    float4 k_cxm = z4, k_cxp = z4, k_cym = z4, k_cyp = z4, k_czp = z4, k_acc = z4, k_gnl = z4, k_r0 = z4;
    float k_czl = 0.f;
    float4 t_m2 = z4, t_m1 = z4;
    float4 cxm = z4;   // -x face coefficients of the plane about to be computed
    if (in_lat && xa - 2 >= 0) {
        const float* q = s.cface + (int64_t)(xa - 2) * plane + e_yz;
        if (vec) cxm = ld4(q); else cxm.x = __ldg(q);
    }

    // the first two U planes (slots 0, 1): own cells only
    mbar_wait_a(fullU, 0u);
    mbar_wait_a(fullU + 8u, 0u);
    float4 u_m1, u_0;
    if (vec) {
        u_m1 = lds128(ringU + so4);
        u_0 = lds128(ringU + slot_b + so4);
    } else {
        u_m1 = z4; u_0 = z4;
        u_m1.x = lds32(ringU + so4);
        u_0.x = lds32(ringU + slot_b + so4);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_a(emptyU);
    uint32_t i1 = 1, i2 = 2, ph2 = 0, it = 0, pht = 0, tw = 0;   // slots: U[p], U[p+1] (+ parity), tables (+ parity), T[p]

#pragma unroll 2
    for (int n = 0; n < n_iter; ++n) {
        mbar_wait_a(fullU + 8u * i2, ph2);
        mbar_wait_a(fullT + 8u * it, pht);
        const uint32_t aC = ringU + i2 * slot_b + so4;      // U[p+1], own cells
        const uint32_t aN = ringU + i1 * slot_b + so4;      // U[p]: y / z neighbours
        const uint32_t aT = ringT + it * (NT * slot_b) + so4;
        const uint32_t aTw = Ts + tw * tplane_b + to4;      // T[p] (written), T[p-1] in the other slot
        const uint32_t aTr = Ts + (tw ^ 1u) * tplane_b + to4;
        float4 n_cxp = z4, n_cym = z4, n_cyp = z4, n_czp = z4, n_acc = z4, n_gnl = z4, n_r0 = z4, u_p1 = z4, t_0 = z4;
        float n_czl = 0.f;
        const bool do_g = is_main && n >= 2 && !(g.dbg & 1);
        if (g.dbg & 1) {
        } else if (vec) {
            u_p1 = lds128(aC);
            const float4 uym = lds128(aN - bw4), uyp = lds128(aN + bw4);
            const float4 di = lds128(aT + T_DI * slot_b), rh = lds128(aT + T_RH * slot_b);
            n_cxp = lds128(aT + T_CX * slot_b);
            n_cyp = lds128(aT + T_CY * slot_b);
            n_cym = lds128(aT + T_CY * slot_b - bw4);
            n_czp = lds128(aT + T_CZ * slot_b);
            float4 kv = z4;
            if (KV) kv = lds128(aT + T_KV * slot_b);
            float4 nla = z4, nlb = z4;
            if (NL) {
                nla = lds128(aT + T_NA * slot_b);
                nlb = lds128(aT + T_NB * slot_b);
            }
            float ul = __shfl_up_sync(vmask, u_0.w, 1), ur = __shfl_down_sync(vmask, u_0.x, 1);
            n_czl = __shfl_up_sync(vmask, n_czp.w, 1);
            if (!left_lane) {
                ul = lds32(aN - 4u);
                n_czl = lds32(aT + T_CZ * slot_b - 4u);
            }
            if (!right_lane) ur = lds32(aN + 16u);
            const float4 czm = make_float4(n_czl, n_czp.x, n_czp.y, n_czp.z);
            const float4 uzm = make_float4(ul, u_0.x, u_0.y, u_0.z), uzp = make_float4(u_0.y, u_0.z, u_0.w, ur);
            float4 rr, dsum;
#define NBM_ST_ROW(c)                                                                                                \
    dsum.c = (((((cxm.c + n_cxp.c) + n_cym.c) + n_cyp.c) + czm.c) + n_czp.c) + kv.c;                                   \
    rr.c = row_val(di.c, rh.c, dsum.c, u_0.c, cxm.c, u_m1.c, n_cxp.c, u_p1.c, n_cym.c, uym.c, n_cyp.c, uyp.c, czm.c,   \
                   uzm.c, n_czp.c, uzp.c);
            NBM_ST_ROW(x) NBM_ST_ROW(y) NBM_ST_ROW(z) NBM_ST_ROW(w)
#undef NBM_ST_ROW
            if (NL) {
                const float4 a = nla, b = nlb;
#define NBM_ST_NL(c)                                                                                                 \
    if (di.c != 0.f) rr.c += a.c * nl_apply(s.nonlinear_m, s.nl_coef_m, u_0.c) + b.c * nl_apply(s.nonlinear_p, s.nl_coef_p, u_0.c); \
    n_gnl.c = nl_dfac(s, a.c, b.c, u_0.c);
                NBM_ST_NL(x) NBM_ST_NL(y) NBM_ST_NL(z) NBM_ST_NL(w)
#undef NBM_ST_NL
            }
            t_0 = tval4(di, rr);
            n_r0 = rr;
            if (active) sts128(aTw, t_0);
            n_acc = make_float4(di.x > 0.f ? dsum.x * t_0.x : (di.x < 0.f ? rr.x : 0.f),
                                di.y > 0.f ? dsum.y * t_0.y : (di.y < 0.f ? rr.y : 0.f),
                                di.z > 0.f ? dsum.z * t_0.z : (di.z < 0.f ? rr.z : 0.f),
                                di.w > 0.f ? dsum.w * t_0.w : (di.w < 0.f ? rr.w : 0.f));
            if (stores && n >= n_r_lo && n < n_r_hi) *reinterpret_cast<float4*>(pR) = rr;
        } else {
            // one cell of a z halo column: only T is needed
            u_p1.x = lds32(aC);
            const float di = lds32(aT + T_DI * slot_b), rh = lds32(aT + T_RH * slot_b);
            const float cxp = lds32(aT + T_CX * slot_b), cyp = lds32(aT + T_CY * slot_b), cym = lds32(aT + T_CY * slot_b - bw4);
            const float czp = lds32(aT + T_CZ * slot_b), czm = lds32(aT + T_CZ * slot_b - 4u);
            const float kv = KV ? lds32(aT + T_KV * slot_b) : 0.f;
            const float dsum = (((((cxm.x + cxp) + cym) + cyp) + czm) + czp) + kv;
            float rr = row_val(di, rh, dsum, u_0.x, cxm.x, u_m1.x, cxp, u_p1.x, cym, lds32(aN - bw4), cyp, lds32(aN + bw4), czm,
                               lds32(aN - 4u), czp, lds32(aN + 4u));
            if (NL) {
                if (di != 0.f)
                    rr += lds32(aT + T_NA * slot_b) * nl_apply(s.nonlinear_m, s.nl_coef_m, u_0.x) +
                          lds32(aT + T_NB * slot_b) * nl_apply(s.nonlinear_p, s.nl_coef_p, u_0.x);
            }
            if (active) sts32(aTw, tval(di, rr));
            n_cxp = make_float4(cxp, 0.f, 0.f, 0.f);
        }
        // G[p-1] on the tile: own T of planes p-2, p-1, p from registers, the y / z neighbours of T[p-1] from shared memory
        // (written before the previous iteration's barrier), coefficients of plane p-1 carried from the previous iteration
        if (do_g) {
            const float4 tym = lds128(aTr - bw4), typ = lds128(aTr + bw4);
            float tl = __shfl_up_sync(0xffffffffu, t_m1.w, 1), tr = __shfl_down_sync(0xffffffffu, t_m1.x, 1);
            if (!left_lane) tl = lds32(aTr - 4u);
            if (!right_lane) tr = lds32(aTr + 16u);
            const float4 tzm = make_float4(tl, t_m1.x, t_m1.y, t_m1.z), tzp = make_float4(t_m1.y, t_m1.z, t_m1.w, tr);
            const float4 czm = make_float4(k_czl, k_czp.x, k_czp.y, k_czp.z);
            float4 gg;
#define NBM_ST_ADJ(c)                                                                                                \
    {                                                                                                                \
        float acc = k_acc.c;                                                                                         \
        acc = fmaf(-k_cxm.c, t_m2.c, acc); acc = fmaf(-k_cxp.c, t_0.c, acc);                                         \
        acc = fmaf(-k_cym.c, tym.c, acc); acc = fmaf(-k_cyp.c, typ.c, acc);                                          \
        acc = fmaf(-czm.c, tzm.c, acc); acc = fmaf(-k_czp.c, tzp.c, acc);                                            \
        gg.c = acc;                                                                                                  \
    }
            NBM_ST_ADJ(x) NBM_ST_ADJ(y) NBM_ST_ADJ(z) NBM_ST_ADJ(w)
#undef NBM_ST_ADJ
            if (NL) {
                gg.x = fmaf(k_gnl.x, k_r0.x, gg.x); gg.y = fmaf(k_gnl.y, k_r0.y, gg.y);
                gg.z = fmaf(k_gnl.z, k_r0.z, gg.z); gg.w = fmaf(k_gnl.w, k_r0.w, gg.w);
            }
            if (stores) *reinterpret_cast<float4*>(pG) = gg;
        }
        // the tables of plane p and the neighbour plane U[p] are consumed; T[p] is complete after the barrier
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_a(emptyT + 8u * it);
            mbar_arrive_a(emptyU + 8u * i1);
        }
        bar_consumers(kConsumers);
        k_cxm = cxm; k_cxp = n_cxp; k_cym = n_cym; k_cyp = n_cyp; k_czp = n_czp; k_czl = n_czl; k_acc = n_acc; k_gnl = n_gnl; k_r0 = n_r0;
        cxm = n_cxp;
        t_m2 = t_m1; t_m1 = t_0;
        u_m1 = u_0; u_0 = u_p1;
        i1 = i2;
        if (++i2 == nsu) { i2 = 0; ph2 ^= 1u; }
        if (++it == nst) { it = 0; pht ^= 1u; }
        tw ^= 1u;
        pR += plane;
        pG += plane;
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
    static EncodeFn fn = []() -> EncodeFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}

static bool encode(CUtensorMap* m, const float* base, const Geom& g) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)g.ez, (cuuint64_t)g.ey, (cuuint64_t)g.ex};
    cuuint64_t strides[2] = {(cuuint64_t)g.ez * 4, (cuuint64_t)g.ez * g.ey * 4};
    cuuint32_t box[3] = {(cuuint32_t)g.bw, (cuuint32_t)g.bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tile shape / x chunks: few CTA waves, little halo, little overhang
static Geom choose_geom(int ex, int ey, int ez, int sms, int n_tab, size_t smem_limit) {
    Geom best{};
    double best_cost = 1e300;
    const int ez4 = ez / 4;
    auto env_i = [](const char* k) { const char* v = getenv(k); return v ? atoi(v) : 0; };
    const int f_gy = env_i("NBM_ST_GY"), f_gzq = env_i("NBM_ST_GZQ"), f_nxc = env_i("NBM_ST_NXC"), f_pf = env_i("NBM_ST_PF");
    const int f_main = env_i("NBM_ST_MAIN"), f_dbg = env_i("NBM_ST_DBG");
    for (int n_main = 128; n_main <= kMainMax; n_main += 128)
    for (int gzq = 4; gzq <= 44; ++gzq) {
        if (f_gzq && gzq != f_gzq) continue;
        if (f_main && n_main != f_main) continue;
        for (int gy = 2; gy <= 44; ++gy) {
            if (f_gy && gy != f_gy) continue;
            const int n_halo = n_main == 256 ? kHaloMax : 64;
            if (gy * gzq > n_main || 2 * (gy + gzq) > n_halo || gy * gzq <= n_main / 2) continue;
            Geom g{};
            g.n_main = n_main; g.n_halo = n_halo; g.dbg = f_dbg;
            g.ex = ex; g.ey = ey; g.ez = ez; g.gy = gy; g.gzq = gzq;
            g.bw = 4 * gzq + 8; g.bh = gy + 4;
            if (g.bw > 256 || g.bh > 256) continue;
            g.slot_floats = ((g.bh * g.bw * 4 + 127) / 128) * 32;
            g.ny_t = (ey + gy - 1) / gy; g.nz_t = (ez4 + gzq - 1) / gzq;
            const int nt = g.ny_t * g.nz_t;
            // deepest prefetch that fits (at least one plane ahead)
            int pf = f_pf ? f_pf : 3;
            for (; pf >= 1; --pf) {
                g.nsu = 2 + pf; g.nst = 1 + pf;
                const size_t b = 256 + 4 * ((size_t)g.slot_floats * (g.nsu + (size_t)g.nst * n_tab) + (size_t)kTRing * (gy + 2) * g.bw);
                if (b <= smem_limit && 2 * (g.nsu + g.nst) <= 16) break;
            }
            if (pf < 1) continue;
            for (int nxc = 1; nxc <= ex; ++nxc) {
                if (f_nxc && nxc != f_nxc) continue;
                const int xchunk = (ex + nxc - 1) / nxc;
                if ((nxc - 1) * xchunk >= ex) continue;   // an empty last chunk
                const int64_t nb = (int64_t)nt * nxc;
                // CTAs per SM: shared memory, and 128 registers per thread of a 64 K file
                const size_t smem_cta = 256 + 4 * ((size_t)g.slot_floats * (g.nsu + (size_t)g.nst * n_tab) + (size_t)kTRing * (gy + 2) * g.bw);
                const int per_sm = (int)fmin(fmin((double)(smem_limit / (smem_cta + 1024)), 512.0 / (n_main + n_halo + 32)), 2.0);
                if (per_sm < 1) continue;
                const int64_t waves = (nb + (int64_t)sms * per_sm - 1) / ((int64_t)sms * per_sm);
                // per plane and CTA: box traffic through L2 (~48 B/clk/SM) against a ~500-clock compute + barrier floor
                const double t_plane = fmax((double)g.bh * g.bw * 4.0 * (n_tab + 1) / 48.0, 500.0) * (pf >= 2 ? 1.0 : 1.15);
                const double cost = (double)waves * (xchunk + 3) * t_plane * (per_sm == 2 ? 1.6 : 1.0);
                if (getenv("NBM_ST_VERBOSE"))
                    fprintf(stderr, "stencil_tma cand: main %d gy %d gzq %d nxc %d pf %d per_sm %d waves %d cost %.0f\n", n_main, gy,
                            gzq, nxc, pf, per_sm, (int)waves, cost);
                if (cost < best_cost) {
                    best_cost = cost;
                    best = g;
                    best.nxc = nxc;
                    best.xchunk = xchunk;
                }
                if (nb > 4 * (int64_t)sms) break;
            }
        }
    }
    return best;
}

struct Cached {
    const void* key[9];
    int dims[3];
    Geom g;
    Maps maps;
    bool ok = false;
};

template <bool KV, bool NL>
static int launch_t(const nbm_shared_step_t& s, int sms, cudaStream_t st) {
    static thread_local Cached cache;
    static unsigned long long configured = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t ne = (int64_t)s.ex * s.ey * s.ez;
    const void* key[9] = {s.U, s.cface, s.dinv, s.rhs, s.kv, s.nl, (const void*)(intptr_t)dev, nullptr, nullptr};
    if (!(cache.ok && memcmp(cache.key, key, sizeof(key)) == 0 && cache.dims[0] == s.ex && cache.dims[1] == s.ey &&
          cache.dims[2] == s.ez)) {
        int smem_max = 0;
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        // 8 KB of the SM's shared memory stay free: the list kernels of the side stream (no shared memory of their own,
        // 1 KB system reservation per CTA) must be able to co-reside with a stencil CTA
        Geom g = choose_geom(s.ex, s.ey, s.ez, sms, n_tables<KV, NL>(), (size_t)smem_max - 8192);
        if (g.gy == 0) {
            set_error("stencil_tma: no tile shape fits");
            return NBM_ERR_UNSUPPORTED;
        }
        if (getenv("NBM_ST_DEBUG"))
            fprintf(stderr, "stencil_tma: lattice %dx%dx%d main %d halo %d tile %dx%d tiles %dx%d nxc %d xchunk %d nsu %d nst %d smem %zu\n",
                    s.ex, s.ey, s.ez, g.n_main, g.n_halo, g.gy, g.gzq, g.ny_t, g.nz_t, g.nxc, g.xchunk, g.nsu, g.nst,
                    smem_bytes<KV, NL>(g));
        Maps& m = cache.maps;
        bool ok = encode(&m.U, s.U, g) && encode(&m.cx, s.cface, g) && encode(&m.cy, s.cface + ne, g) &&
                  encode(&m.cz, s.cface + 2 * ne, g) && encode(&m.dinv, s.dinv, g) && encode(&m.rhs, s.rhs, g);
        if (KV) ok = ok && encode(&m.kv, s.kv, g);
        if (NL) ok = ok && encode(&m.nla, s.nl, g) && encode(&m.nlb, s.nl + ne, g);
        if (!ok) {
            set_error("stencil_tma: cuTensorMapEncodeTiled failed");
            return NBM_ERR_CUDA;
        }
        memcpy(cache.key, key, sizeof(key));
        cache.dims[0] = s.ex; cache.dims[1] = s.ey; cache.dims[2] = s.ez;
        cache.g = g;
        cache.ok = true;
    }
    const Geom& g = cache.g;
    if (!(configured & (1ull << (dev & 63)))) {
        int smem_max = 0;
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaError_t e = cudaFuncSetAttribute(stencil_tma_kernel<KV, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(stencil_tma_kernel<KV, NL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return cuda_check(e, "stencil_tma attribute");
        configured |= 1ull << (dev & 63);
    }
    const int grid = g.ny_t * g.nz_t * g.nxc;
    cudaError_t le = launch_pdl(stencil_tma_kernel<KV, NL>, dim3(grid), dim3(g.n_main + g.n_halo + 32), smem_bytes<KV, NL>(g), st,
                                cache.maps, g, s);
    return le == cudaSuccess ? 0 : cuda_check(le, "stencil_tma launch");
}

// applicable: faces table, 16-byte lattice rows and array bases
static bool applicable(const nbm_shared_step_t& s) {
    if (!s.faces || !s.cface || !s.dinv || (s.ez % 4) != 0) return false;
    if ((((uintptr_t)s.U | (uintptr_t)s.R | (uintptr_t)s.G | (uintptr_t)s.cface | (uintptr_t)s.dinv | (uintptr_t)s.rhs |
          (uintptr_t)s.kv | (uintptr_t)s.nl) & 15) != 0)
        return false;
    return encode_fn() != nullptr;
}

static int launch(const nbm_shared_step_t& s, int sms, cudaStream_t st) {
    if (s.kv) return s.nl ? launch_t<true, true>(s, sms, st) : launch_t<true, false>(s, sms, st);
    return s.nl ? launch_t<false, true>(s, sms, st) : launch_t<false, false>(s, sms, st);
}

}  // namespace stencil_tma
