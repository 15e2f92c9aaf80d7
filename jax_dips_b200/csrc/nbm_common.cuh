// Shared device helpers: level-set interpolation on the ghosted lvl grid, lattice addressing,
// error plumbing.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nbm_b200.h"

namespace nbm {

void set_error(const char* fmt, ...);
int cuda_check(cudaError_t e, const char* what);

#define NBM_REQUIRE(cond, msg)                          \
    do {                                                \
        if (!(cond)) {                                  \
            nbm::set_error("%s: %s", __func__, msg);    \
            return NBM_ERR_BAD_ARG;                     \
        }                                               \
    } while (0)

#define NBM_LAUNCH_CHECK(what)                                              \
    do {                                                                    \
        int _rc = nbm::cuda_check(cudaGetLastError(), what);                \
        if (_rc) return _rc;                                                \
    } while (0)

static inline cudaStream_t as_stream(nbm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched with `launch_pdl` may have its CTAs placed while the previous
// kernel of the stream is still draining; it must call pdl_wait() before it touches global memory (the wait returns
// when the previous kernel has completed and its writes are visible).  pdl_trigger() lets the NEXT kernel start that
// early placement.  Without a programmatic edge both are no-ops.  The launch attribute is opt-in (NBM_PDL=1): it measured no gain, see pdl_enabled().
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// a14: phi from the ghosted lvl grid.  Follows interpolate.py:946-956 (cell index with the
// `i<=1 -> 2` clamp, so the first interior cell extrapolates from its neighbour), :1001-1009
// (trilinear) and :485-565 (non-oscillatory quadratic correction, un-normalised second
// differences, min over the 8 cell corners).  geometry/level_set.py:34-48 for the perturbation.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lvl_cell(float p, const float* __restrict__ a, int n) {
    float a0 = __ldg(a), a1 = __ldg(a + 1);
    float d = a1 - a0;
    int i = (int)__fdiv_rn(p - a0, d);  // astype(int32): truncation toward zero
    if (i >= n - 1) i = n - 2;
    if (i <= 1) i = 2;
    return i;
}

__device__ __forceinline__ float lvl_at(const nbm_lvl_t& L, int i, int j, int k) {
    i = min(max(i, 0), L.gx - 1);  // XLA gather clamps out-of-range indices (only i+2 at the last cell)
    j = min(max(j, 0), L.gy - 1);
    k = min(max(k, 0), L.gz - 1);
    return __ldg(L.phi_g + ((size_t)i * L.gy + j) * L.gz + k);
}

__device__ __forceinline__ float phi_at(const nbm_lvl_t& L, float xp, float yp, float zp) {
    int i = lvl_cell(xp, L.xg, L.gx), j = lvl_cell(yp, L.yg, L.gy), k = lvl_cell(zp, L.zg, L.gz);
    float x0 = __ldg(L.xg + i), x1 = __ldg(L.xg + i + 1);
    float y0 = __ldg(L.yg + j), y1 = __ldg(L.yg + j + 1);
    float z0 = __ldg(L.zg + k), z1 = __ldg(L.zg + k + 1);
    float xd = __fdiv_rn(xp - x0, x1 - x0);
    float yd = __fdiv_rn(yp - y0, y1 - y0);
    float zd = __fdiv_rn(zp - z0, z1 - z0);
    const float* b = L.phi_g + ((size_t)i * L.gy + j) * L.gz + k;
    size_t sx = (size_t)L.gy * L.gz, sy = L.gz;
    float c000 = __ldg(b), c001 = __ldg(b + 1);
    float c010 = __ldg(b + sy), c011 = __ldg(b + sy + 1);
    float c100 = __ldg(b + sx), c101 = __ldg(b + sx + 1);
    float c110 = __ldg(b + sx + sy), c111 = __ldg(b + sx + sy + 1);
    float c00 = c000 * (1.0f - xd) + c100 * xd;
    float c01 = c001 * (1.0f - xd) + c101 * xd;
    float c10 = c010 * (1.0f - xd) + c110 * xd;
    float c11 = c011 * (1.0f - xd) + c111 * xd;
    float c0 = c00 * (1.0f - yd) + c10 * yd;
    float c1 = c01 * (1.0f - yd) + c11 * yd;
    float c = c0 * (1.0f - zd) + c1 * zd;
    if (L.interp == NBM_INTERP_QUADRATIC) {
        float mx = 3.4e38f, my = 3.4e38f, mz = 3.4e38f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    int ii = i + a, jj = j + bb, kk = k + cc;
                    float ctr = lvl_at(L, ii, jj, kk);
                    mx = fminf(mx, fabsf(lvl_at(L, ii + 1, jj, kk) - 2.0f * ctr + lvl_at(L, ii - 1, jj, kk)));
                    my = fminf(my, fabsf(lvl_at(L, ii, jj + 1, kk) - 2.0f * ctr + lvl_at(L, ii, jj - 1, kk)));
                    mz = fminf(mz, fabsf(lvl_at(L, ii, jj, kk + 1) - 2.0f * ctr + lvl_at(L, ii, jj, kk - 1)));
                }
        c = c - mx * 0.5f * xd * (1.0f - xd) - my * 0.5f * yd * (1.0f - yd) - mz * 0.5f * zd * (1.0f - zd);
    }
    if (L.perturb_eps != 0.0f) c = c + (c > 0.0f ? L.perturb_eps : -L.perturb_eps);  // sign_pm: 0 -> -1
    return c;
}

// site id -> position.  sid = k*n + e, e = (ix*ny + iy)*nz + iz
struct SitePos {
    float x, y, z;
    int k, ix, iy, iz;
};

__device__ __forceinline__ SitePos site_position(const nbm_lattice_t& lat, int64_t sid) {
    int64_t n = (int64_t)lat.nx * lat.ny * lat.nz;
    SitePos s;
    s.k = (int)(sid / n);
    int64_t e = sid - (int64_t)s.k * n;
    s.iz = (int)(e % lat.nz);
    int64_t t = e / lat.nz;
    s.iy = (int)(t % lat.ny);
    s.ix = (int)(t / lat.ny);
    s.x = __ldg(lat.xs + s.ix) + lat.shift[s.k][0];
    s.y = __ldg(lat.ys + s.iy) + lat.shift[s.k][1];
    s.z = __ldg(lat.zs + s.iz) + lat.shift[s.k][2];
    return s;
}

}  // namespace nbm
