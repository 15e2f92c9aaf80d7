// Per-level (parameter-free) kernels: ghost layer, level-set interpolation, site classification,
// cut-cell geometry (K1), regression extrapolation weights (K2a/K2b), row assembly (K2c).
// Compiled with -fmad=false so that the fp32 sign / ordering decisions (crossing flags, tet cases,
// vertex sorting, on-face tests) see exactly the values a plain fp32 evaluation of the reference
// formulas produces.
#include <cub/cub.cuh>
#include <stdarg.h>
#include <string.h>
#include "nbm_common.cuh"

namespace nbm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return NBM_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NBM_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// ghost layer (interpolate.py:762-816): x faces from the interior, then y on the x-ghosted array,
// then z on the xy-ghosted array.
// ---------------------------------------------------------------------------------------------
__global__ void ghost_fill_interior(const float* __restrict__ phi, float* __restrict__ g, int nx, int ny, int nz) {
    int64_t n = (int64_t)nx * ny * nz;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int k = (int)(e % nz);
    int64_t t = e / nz;
    int j = (int)(t % ny), i = (int)(t / ny);
    int gy = ny + 2, gz = nz + 2;
    g[((size_t)(i + 1) * gy + (j + 1)) * gz + (k + 1)] = phi[e];
}

// axis 0: x ghost planes over interior (j,k); axis 1: y planes over all i, interior k...; the
// reference applies y to the full [:, *, :] slab (including x ghosts, and the still-zero z ghosts)
// and finally z to the full [:, :, *] slab, which overwrites every earlier z-ghost value.
__global__ void ghost_extrapolate(float* __restrict__ g, int gx, int gy, int gz, int axis) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t sx = (size_t)gy * gz, sy = gz;
    if (axis == 0) {
        int64_t n = (int64_t)(gy - 2) * (gz - 2);
        if (idx >= n) return;
        int k = (int)(idx % (gz - 2)) + 1, j = (int)(idx / (gz - 2)) + 1;
        size_t o = j * sy + k;
        g[o] = 2.0f * g[o + sx] - g[o + 2 * sx];
        g[o + (gx - 1) * sx] = 2.0f * g[o + (gx - 2) * sx] - g[o + (gx - 3) * sx];
    } else if (axis == 1) {
        int64_t n = (int64_t)gx * gz;
        if (idx >= n) return;
        int k = (int)(idx % gz), i = (int)(idx / gz);
        size_t o = i * sx + k;
        g[o] = 2.0f * g[o + sy] - g[o + 2 * sy];
        g[o + (gy - 1) * sy] = 2.0f * g[o + (gy - 2) * sy] - g[o + (gy - 3) * sy];
    } else {
        int64_t n = (int64_t)gx * gy;
        if (idx >= n) return;
        size_t o = (size_t)idx * gz;
        g[o] = 2.0f * g[o + 1] - g[o + 2];
        g[o + gz - 1] = 2.0f * g[o + gz - 2] - g[o + gz - 3];
    }
}

__global__ void ghost_coords(const float* __restrict__ a, float* __restrict__ ag, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n + 1) return;
    float v;
    if (i == 0) v = a[0] - (a[1] - a[0]);
    else if (i == n + 1) v = a[n - 1] + (a[n - 1] - a[n - 2]);
    else v = a[i - 1];
    ag[i] = v;
}

__global__ void phi_interp_kernel(nbm_lvl_t L, const float* __restrict__ pts, int64_t n, float* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    out[e] = phi_at(L, pts[3 * e], pts[3 * e + 1], pts[3 * e + 2]);
}

// ---------------------------------------------------------------------------------------------
// classification (geometric_integrations_per_point.py:203-263)
// corner order 000,100,101,001,010,110,011,111 (:212-224); flag = (#neg in {0,8}) ? sign(phi_000) : 0
// ---------------------------------------------------------------------------------------------
__constant__ float c_corner_sign[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, -1, 1}, {-1, -1, 1},
                                          {-1, 1, -1},  {1, 1, -1},  {-1, 1, 1}, {1, 1, 1}};

__device__ __forceinline__ void corner_phis(const nbm_lvl_t& L, float x, float y, float z, float dx, float dy,
                                            float dz, float (&P)[8][3], float (&ph)[8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        P[c][0] = (c_corner_sign[c][0] * dx) * 0.5f + x;
        P[c][1] = (c_corner_sign[c][1] * dy) * 0.5f + y;
        P[c][2] = (c_corner_sign[c][2] * dz) * 0.5f + z;
        ph[c] = phi_at(L, P[c][0], P[c][1], P[c][2]);
    }
}

__global__ void classify_kernel(nbm_lvl_t L, nbm_lattice_t lat, float dx, float dy, float dz,
                                int8_t* __restrict__ flag, uint8_t* __restrict__ side, int64_t total) {
    int64_t sid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sid >= total) return;
    SitePos s = site_position(lat, sid);
    float ph0 = phi_at(L, s.x, s.y, s.z);
    side[sid] = (uint8_t)((ph0 >= 0.0f ? 1 : 0) | (ph0 > 0.0f ? 2 : 0));
    bool is_site = s.ix >= lat.lo[0] && s.ix < lat.hi[0] && s.iy >= lat.lo[1] && s.iy < lat.hi[1] &&
                   s.iz >= lat.lo[2] && s.iz < lat.hi[2];
    if (!is_site) {
        flag[sid] = 2;
        return;
    }
    float P[8][3], ph[8];
    corner_phis(L, s.x, s.y, s.z, dx, dy, dz, P, ph);
    int neg = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) neg += (ph[c] < 0.0f) ? 1 : 0;
    int8_t f = 0;
    if (neg == 0 || neg == 8) f = (ph[0] > 0.0f) ? 1 : ((ph[0] < 0.0f) ? -1 : 0);
    flag[sid] = f;
}

// ---------------------------------------------------------------------------------------------
// compaction of crossed sites
// ---------------------------------------------------------------------------------------------
struct IsZero {
    __host__ __device__ bool operator()(const int8_t& v) const { return v == 0; }
};

__global__ void cidx_fill(int32_t* cidx, int64_t n) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) cidx[e] = -1;
}

__global__ void cidx_scatter(const int64_t* __restrict__ idx, const int64_t* __restrict__ count, int64_t capacity,
                             int32_t* __restrict__ cidx) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = min(*count, capacity);
    if (c < n) cidx[idx[c]] = (int32_t)c;
}

// ---------------------------------------------------------------------------------------------
// K1: cut cell (Min & Gibou 2007 middle-cut triangulation as the reference applies it)
// ---------------------------------------------------------------------------------------------
__constant__ int c_tets[5][4] = {{0, 1, 4, 3}, {5, 1, 4, 7}, {2, 1, 7, 3}, {6, 7, 4, 3}, {7, 1, 4, 3}};  // :312-316
// faces owned by each tet (:841-852): face ids 0..5 = x-,x+,y-,y+,z-,z+ ; -1 = none
__constant__ int c_tet_faces[5][3] = {{0, 2, 4}, {1, 3, 4}, {1, 2, 5}, {0, 3, 5}, {-1, -1, -1}};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float comp(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// exact-arithmetic equal of sqrt|det(E E^T)|/6 (:370-374) without squaring the condition number
__device__ __forceinline__ float tet_volume(V3 a, V3 b, V3 c, V3 d) {
    V3 e1 = b - a, e2 = c - a, e3 = d - a;
    float det = e1.x * (e2.y * e3.z - e2.z * e3.y) - e1.y * (e2.x * e3.z - e2.z * e3.x) +
                e1.z * (e2.x * e3.y - e2.y * e3.x);
    return fabsf(det) * (1.0f / 6.0f);
}
// equal of 0.5*sqrt|det(E E^T)| (:377-380)
__device__ __forceinline__ float tri_area(V3 a, V3 b, V3 c) {
    V3 e1 = b - a, e2 = c - a;
    float cx = e1.y * e2.z - e1.z * e2.y, cy = e1.z * e2.x - e1.x * e2.z, cz = e1.x * e2.y - e1.y * e2.x;
    return 0.5f * sqrtf(cx * cx + cy * cy + cz * cz);
}

// stable ascending sort of 4 keys (jnp.argsort is stable)
__device__ __forceinline__ void sort4(float (&key)[4], V3 (&v)[4]) {
#pragma unroll
    for (int i = 1; i < 4; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j) {
            if (key[j - 1] > key[j]) {
                float tk = key[j]; key[j] = key[j - 1]; key[j - 1] = tk;
                V3 tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv;
            }
        }
    }
}

// (phi_a S_b - phi_b S_a)/(phi_a - phi_b)  (:70)
__device__ __forceinline__ V3 cut(const float (&ph)[4], const V3 (&S)[4], int a, int b) {
    float den = ph[a] - ph[b];
    return V3{__fdiv_rn(ph[a] * S[b].x - ph[b] * S[a].x, den), __fdiv_rn(ph[a] * S[b].y - ph[b] * S[a].y, den),
              __fdiv_rn(ph[a] * S[b].z - ph[b] * S[a].z, den)};
}

__device__ __forceinline__ bool on_face(float c, float face, float atol) {
    return fabsf(c - face) <= atol + 1e-5f * fabsf(face);  // jnp.isclose(c, face, rtol=1e-5, atol=atol)
}

__global__ void cutcell_kernel(nbm_lvl_t L, nbm_lattice_t lat, float dx, float dy, float dz,
                               const int64_t* __restrict__ idx, int64_t n, float* __restrict__ frac,
                               float* __restrict__ tri, float* __restrict__ tri_area_out) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    SitePos s = site_position(lat, idx[c]);
    float P[8][3], ph8[8];
    if (L.corner_phi) {   // sampled level set: positions formed as in corner_phis, values handed in by the host
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            P[v][0] = (c_corner_sign[v][0] * dx) * 0.5f + s.x;
            P[v][1] = (c_corner_sign[v][1] * dy) * 0.5f + s.y;
            P[v][2] = (c_corner_sign[v][2] * dz) * 0.5f + s.z;
            ph8[v] = L.corner_phi[c * 8 + v];
        }
    } else {
        corner_phis(L, s.x, s.y, s.z, dx, dy, dz, P, ph8);
    }
    const float face_coord[6] = {s.x - 0.5f * dx, s.x + 0.5f * dx, s.y - 0.5f * dy,
                                 s.y + 0.5f * dy, s.z - 0.5f * dz, s.z + 0.5f * dz};
    const float face_atol[6] = {1e-10f * dx, 1e-10f * dx, 1e-10f * dy, 1e-10f * dy, 1e-10f * dz, 1e-10f * dz};
    float area_m[6] = {0, 0, 0, 0, 0, 0};
    float vol_m = 0.0f;
    float* tri_c = tri + c * 90;
    float* ta_c = tri_area_out + c * 10;

    for (int t = 0; t < 5; ++t) {
        V3 S[4], Ss[4];
        float ph[4], ps[4];
        int eta = 0;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            int ci = c_tets[t][v];
            S[v] = v3(P[ci][0], P[ci][1], P[ci][2]);
            ph[v] = ph8[ci];
            eta += (ph[v] < 0.0f) ? 1 : 0;
            Ss[v] = S[v];
            ps[v] = ph[v];
        }
        sort4(ps, Ss);

        // ---- Gamma pieces (:53-107)
        V3 g[2][3];
        int ng = 0;
        if (eta == 1) {
            g[0][0] = cut(ps, Ss, 0, 1); g[0][1] = cut(ps, Ss, 0, 2); g[0][2] = cut(ps, Ss, 0, 3);
            ng = 1;
        } else if (eta == 2) {
            V3 Q0 = cut(ps, Ss, 0, 2), Q1 = cut(ps, Ss, 0, 3), Q2 = cut(ps, Ss, 1, 3), Q5 = cut(ps, Ss, 1, 2);
            g[0][0] = Q0; g[0][1] = Q1; g[0][2] = Q2;
            g[1][0] = Q0; g[1][1] = Q5; g[1][2] = Q2;
            ng = 2;
        } else if (eta == 3) {
            // eta_3: eta_1 applied to -phi, re-sorted (stable) (:76-78)
            V3 Sn[4];
            float pn[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) { Sn[v] = S[v]; pn[v] = -1.0f * ph[v]; }
            sort4(pn, Sn);
            g[0][0] = cut(pn, Sn, 0, 1); g[0][1] = cut(pn, Sn, 0, 2); g[0][2] = cut(pn, Sn, 0, 3);
            ng = 1;
        }
        for (int j = 0; j < 2; ++j) {
            float a = 0.0f;
            if (j < ng) {
                a = tri_area(g[j][0], g[j][1], g[j][2]);
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    tri_c[(t * 2 + j) * 9 + v * 3 + 0] = g[j][v].x;
                    tri_c[(t * 2 + j) * 9 + v * 3 + 1] = g[j][v].y;
                    tri_c[(t * 2 + j) * 9 + v * 3 + 2] = g[j][v].z;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 9; ++q) tri_c[(t * 2 + j) * 9 + q] = 0.0f;
            }
            ta_c[t * 2 + j] = a;
        }

        // ---- Omega^- pieces (:111-198)
        V3 om[3][4];
        int no = 0;
        if (eta == 1) {
            om[0][0] = Ss[0]; om[0][1] = cut(ps, Ss, 0, 1); om[0][2] = cut(ps, Ss, 0, 2); om[0][3] = cut(ps, Ss, 0, 3);
            no = 1;
        } else if (eta == 2) {
            V3 Q0 = Ss[0], Q1 = Ss[1], Q2 = cut(ps, Ss, 0, 2), Q3 = cut(ps, Ss, 1, 3), Q4 = cut(ps, Ss, 1, 2),
               Q5 = cut(ps, Ss, 0, 3);
            om[0][0] = Q0; om[0][1] = Q1; om[0][2] = Q2; om[0][3] = Q3;
            om[1][0] = Q4; om[1][1] = Q1; om[1][2] = Q2; om[1][3] = Q3;
            om[2][0] = Q0; om[2][1] = Q5; om[2][2] = Q2; om[2][3] = Q3;
            no = 3;
        } else if (eta == 3) {
            V3 Q0 = Ss[0], Q1 = Ss[1], Q2 = Ss[2], Q3 = cut(ps, Ss, 1, 3), Q4 = cut(ps, Ss, 0, 3),
               Q5 = cut(ps, Ss, 2, 3);
            om[0][0] = Q0; om[0][1] = Q1; om[0][2] = Q2; om[0][3] = Q3;
            om[1][0] = Q0; om[1][1] = Q4; om[1][2] = Q2; om[1][3] = Q3;
            om[2][0] = Q5; om[2][1] = Q4; om[2][2] = Q2; om[2][3] = Q3;
            no = 3;
        } else if (eta == 4) {
            om[0][0] = S[0]; om[0][1] = S[1]; om[0][2] = S[2]; om[0][3] = S[3];
            no = 1;
        }
        for (int j = 0; j < no; ++j) {
            vol_m += tet_volume(om[j][0], om[j][1], om[j][2], om[j][3]);
            // face areas (:584-839): a piece contributes the triangle of its on-face vertices iff
            // exactly 3 of its 4 vertices are on the face
#pragma unroll
            for (int ff = 0; ff < 3; ++ff) {
                int f = c_tet_faces[t][ff];
                if (f < 0) continue;
                int axis = f >> 1;
                V3 onv[4];
                int cnt = 0;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (on_face(comp(om[j][v], axis), face_coord[f], face_atol[f])) {
                        onv[cnt < 3 ? cnt : 3] = om[j][v];
                        ++cnt;
                    }
                }
                if (cnt == 3) area_m[f] += tri_area(onv[0], onv[1], onv[2]);
            }
        }
    }
    // zero padded pieces of the reference sit at the origin: they are on a face only if the face
    // coordinate is (numerically) 0, then all 4 vertices are on it -> count 4 -> no area.  Nothing to add.
    float vol = dx * dy * dz;
    float nominal[6] = {dy * dz, dy * dz, dx * dz, dx * dz, dx * dy, dx * dy};
    float* fr = frac + c * 14;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        float ap = nominal[f] - area_m[f];
        fr[2 * f] = area_m[f];
        fr[2 * f + 1] = ap < 0.0f ? 0.0f : ap;
    }
    float vp = vol - vol_m;
    fr[12] = vol_m;
    fr[13] = vp < 0.0f ? 0.0f : vp;
}

// ---------------------------------------------------------------------------------------------
// K2a: regression geometry (discretization.py:238-296)
// ---------------------------------------------------------------------------------------------
// symmetric 3x3 pseudo-inverse with jnp.linalg.pinv's cutoff (rcond = 10*max(M,N)*eps_f32):
// cyclic Jacobi eigen-decomposition, eigenvalues <= rcond*max dropped.
__device__ void pinv_sym3(const float (&A)[6], float (&Pinv)[6]) {
    // A = [a00 a01 a02 a11 a12 a22]
    float a[3][3] = {{A[0], A[1], A[2]}, {A[1], A[3], A[4]}, {A[2], A[4], A[5]}};
    float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        float off = fabsf(a[0][1]) + fabsf(a[0][2]) + fabsf(a[1][2]);
        float dia = fabsf(a[0][0]) + fabsf(a[1][1]) + fabsf(a[2][2]);
        if (off <= 1e-12f * dia || off == 0.0f) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                float apq = a[p][q];
                if (apq == 0.0f) continue;
                float theta = (a[q][q] - a[p][p]) / (2.0f * apq);
                float t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
                float cs = 1.0f / sqrtf(t * t + 1.0f), sn = t * cs;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    float akp = a[k][p], akq = a[k][q];
                    a[k][p] = cs * akp - sn * akq;
                    a[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    float apk = a[p][k], aqk = a[q][k];
                    a[p][k] = cs * apk - sn * aqk;
                    a[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    float vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    float lam[3] = {a[0][0], a[1][1], a[2][2]};
    float lmax = fmaxf(fabsf(lam[0]), fmaxf(fabsf(lam[1]), fabsf(lam[2])));
    const float rcond = 10.0f * 3.0f * 1.1920929e-07f;
    float inv[3];
    for (int i = 0; i < 3; ++i) inv[i] = (fabsf(lam[i]) > rcond * lmax) ? 1.0f / lam[i] : 0.0f;
    int o = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            Pinv[o++] = V[i][0] * inv[0] * V[j][0] + V[i][1] * inv[1] * V[j][1] + V[i][2] * inv[2] * V[j][2];
}

__global__ void regression_kernel(nbm_lvl_t L, nbm_lattice_t lat, float dx, float dy, float dz,
                                  const int64_t* __restrict__ idx, int64_t n, float* __restrict__ pos,
                                  float* __restrict__ proj, float* __restrict__ delta, float* __restrict__ Cm,
                                  float* __restrict__ Cp, uint32_t* __restrict__ cube_side) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    SitePos s = site_position(lat, idx[c]);
    float Ap[6] = {0, 0, 0, 0, 0, 0}, Am[6] = {0, 0, 0, 0, 0, 0};
    uint32_t wp_bits = 0, wm_bits = 0, side_bits = 0;
    for (int q = 0; q < 27; ++q) {
        float X0 = (float)(q % 3 - 1) * dx, X1 = (float)((q / 3) % 3 - 1) * dy, X2 = (float)(q / 9 - 1) * dz;
        float ph = L.cube_phi ? L.cube_phi[c * 27 + q] : phi_at(L, s.x + X0, s.y + X1, s.z + X2);
        if (ph >= 0.0f) side_bits |= (1u << q);
        if (ph > 0.0f) {
            wp_bits |= (1u << q);
            Ap[0] += X0 * X0; Ap[1] += X0 * X1; Ap[2] += X0 * X2; Ap[3] += X1 * X1; Ap[4] += X1 * X2; Ap[5] += X2 * X2;
        } else if (ph < 0.0f) {
            wm_bits |= (1u << q);
            Am[0] += X0 * X0; Am[1] += X0 * X1; Am[2] += X0 * X2; Am[3] += X1 * X1; Am[4] += X1 * X2; Am[5] += X2 * X2;
        }
    }
    // normal (:199-218)
    float gx, gy, gz, d0;
    if (L.cube_phi) {   // s +- d e_a are the cube vertices 13 +- 1, 13 +- 3, 13 +- 9 (same fp32 positions)
        const float* cp = L.cube_phi + c * 27;
        gx = (cp[14] - cp[12]) / (2.0f * dx);
        gy = (cp[16] - cp[10]) / (2.0f * dy);
        gz = (cp[22] - cp[4]) / (2.0f * dz);
        d0 = cp[13];
    } else {
        gx = (phi_at(L, s.x + dx, s.y, s.z) - phi_at(L, s.x - dx, s.y, s.z)) / (2.0f * dx);
        gy = (phi_at(L, s.x, s.y + dy, s.z) - phi_at(L, s.x, s.y - dy, s.z)) / (2.0f * dy);
        gz = (phi_at(L, s.x, s.y, s.z + dz) - phi_at(L, s.x, s.y, s.z - dz)) / (2.0f * dz);
        d0 = phi_at(L, s.x, s.y, s.z);
    }
    float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
    float n0 = gx / nrm, n1 = gy / nrm, n2 = gz / nrm;
    float Pp[6], Pm[6];
    pinv_sym3(Ap, Pp);
    pinv_sym3(Am, Pm);
    // C_q = n . D[:,q] = w_q * (pinv n) . X_q   (pinv symmetric)
    float vp0 = Pp[0] * n0 + Pp[1] * n1 + Pp[2] * n2, vp1 = Pp[1] * n0 + Pp[3] * n1 + Pp[4] * n2,
          vp2 = Pp[2] * n0 + Pp[4] * n1 + Pp[5] * n2;
    float vm0 = Pm[0] * n0 + Pm[1] * n1 + Pm[2] * n2, vm1 = Pm[1] * n0 + Pm[3] * n1 + Pm[4] * n2,
          vm2 = Pm[2] * n0 + Pm[4] * n1 + Pm[5] * n2;
    for (int q = 0; q < 27; ++q) {
        float X0 = (float)(q % 3 - 1) * dx, X1 = (float)((q / 3) % 3 - 1) * dy, X2 = (float)(q / 9 - 1) * dz;
        float cp = (wp_bits >> q & 1u) ? (vp0 * X0 + vp1 * X1 + vp2 * X2) : 0.0f;
        float cm = (wm_bits >> q & 1u) ? (vm0 * X0 + vm1 * X1 + vm2 * X2) : 0.0f;
        Cp[c * 27 + q] = isfinite(cp) ? cp : 0.0f;
        Cm[c * 27 + q] = isfinite(cm) ? cm : 0.0f;
    }
    pos[c * 3 + 0] = s.x; pos[c * 3 + 1] = s.y; pos[c * 3 + 2] = s.z;
    proj[c * 3 + 0] = s.x - d0 * n0; proj[c * 3 + 1] = s.y - d0 * n1; proj[c * 3 + 2] = s.z - d0 * n2;
    delta[c] = d0;
    cube_side[c] = side_bits;
}

// ---------------------------------------------------------------------------------------------
// K2b: jump weights (discretization.py:268-284 zeta/gamma, :464-513 the four extrapolations)
// ---------------------------------------------------------------------------------------------
__global__ void site_weights_kernel(int64_t n, const float* __restrict__ delta, const float* __restrict__ Cm,
                                    const float* __restrict__ Cp, const float* __restrict__ mu_m_s,
                                    const float* __restrict__ mu_p_s, const float* __restrict__ alpha_proj,
                                    const float* __restrict__ beta_proj, const float* __restrict__ mu_m_proj,
                                    const float* __restrict__ mu_p_proj, float* __restrict__ B) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float d = delta[c], mum = mu_m_s[c], mup = mu_p_s[c];
    float alpha = alpha_proj[c], beta = beta_proj[c];
    bool minus_bigger = mum > mup;   // :511-513
    bool plus_side = d > 0.0f;       // node sits in Omega^+ -> u^- is extrapolated
    // which coefficient family: (minus_bigger, plus_side) -> gamma_m ; (minus_bigger, !plus) -> zeta_m ;
    // (!minus_bigger, plus) -> zeta_p ; (!minus_bigger, !plus) -> gamma_p
    bool use_m = minus_bigger;
    bool use_gamma = (minus_bigger == plus_side);
    const float* C = use_m ? (Cm + c * 27) : (Cp + c * 27);
    float fac = use_m ? ((mup - mum) / mup) * d : ((mup - mum) / mum) * d;
    float z[27];
    float zsum = 0.0f;
    for (int q = 0; q < 27; ++q) { z[q] = fac * C[q]; zsum += z[q]; }
    float zeta = (zsum - z[13]) * -1.0f;
    float coef[27];
    float cS;  // the scalar zeta_ijk or gamma_ijk
    if (use_gamma) {
        float den = use_m ? (1.0f - zeta) : (1.0f + zeta);
        float gsum = 0.0f;
        for (int q = 0; q < 27; ++q) { coef[q] = z[q] / den; gsum += coef[q]; }
        cS = (gsum - coef[13]) * -1.0f;
    } else {
        for (int q = 0; q < 27; ++q) coef[q] = z[q];
        cS = zeta;
    }
    float jump = alpha + d * (beta / (minus_bigger ? mu_p_proj[c] : mu_m_proj[c]));
    // E = -coef . U + (1 - cS + coef13) u + r
    float r;
    if (minus_bigger) r = plus_side ? (-1.0f * (1.0f - cS) * jump) : jump;
    else r = plus_side ? (-1.0f * jump) : ((1.0f - cS) * jump);
    float* Bc = B + c * 28;
    for (int q = 0; q < 27; ++q) Bc[q] = -coef[q];
    Bc[13] += (1.0f - cS + coef[13]);
    Bc[27] = r;
}

// ---------------------------------------------------------------------------------------------
// K2c: row assembly (discretization.py:335-420)
// ---------------------------------------------------------------------------------------------
__constant__ int c_slot_off[7][3] = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};

__global__ void assemble_kernel(nbm_assemble_t a) {
    int64_t np = (int64_t)a.pts.nx * a.pts.ny * a.pts.nz;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    int k = (int)(p % a.pts.nz);
    int64_t t = p / a.pts.nz;
    int j = (int)(t % a.pts.ny), i = (int)(t / a.pts.ny);
    float x = a.pts.xs[i], y = a.pts.ys[j], z = a.pts.zs[k];
    float dx = a.dx, dy = a.dy, dz = a.dz;
    int64_t out = a.out_off + i * a.out_stride[0] + j * a.out_stride[1] + k * a.out_stride[2];

    // box boundary (:319-333)
    bool bnd = fabsf(x - a.bounds[0]) < 1e-6f * dx || fabsf(x - a.bounds[1]) < 1e-6f * dx ||
               fabsf(y - a.bounds[2]) < 1e-6f * dy || fabsf(y - a.bounds[3]) < 1e-6f * dy ||
               fabsf(z - a.bounds[4]) < 1e-6f * dz || fabsf(z - a.bounds[5]) < 1e-6f * dz;
    const bool bnd_early = bnd && !a.faces;   // row is u = g; only the preconditioner input is still needed
    if (bnd_early) {
        // lhs = u*vol, diag = vol, rhs = g*vol (:389-393, :410-411)
        a.w[out] = 1.0f;
        for (int s = 1; s < 7; ++s) a.w[s * a.n_out + out] = 0.0f;
        a.rhs[out] = a.g_dir[p];
        if (a.nl) { a.nl[out] = 0.0f; a.nl[a.n_out + out] = 0.0f; }
        a.irr[out] = -1;
        if (!a.coef26) return;
    }
    // site ids of the 7 stencil slots
    int64_t sid[7];
    for (int s = 0; s < 7; ++s) {
        if (a.shared)
            sid[s] = ((int64_t)(i + a.pt_off[0] + c_slot_off[s][0]) * a.site_dims[1] +
                      (j + a.pt_off[1] + c_slot_off[s][1])) * a.site_dims[2] + (k + a.pt_off[2] + c_slot_off[s][2]);
        else
            sid[s] = (int64_t)s * np + p;
    }
    int8_t f0 = a.flag[sid[0]];
    float vol = dx * dy * dz;
    float nominal[6] = {dy * dz, dy * dz, dx * dz, dx * dz, dx * dy, dx * dy};
    float dd[6] = {dx, dx, dy, dy, dz, dz};
    float am[6], ap[6], Vm, Vp, bg = 0.0f;
    if (f0 == 0) {
        int32_t c = a.cidx[sid[0]];
        const float* fr = a.frac + (int64_t)c * 14;
        for (int f = 0; f < 6; ++f) { am[f] = fr[2 * f]; ap[f] = fr[2 * f + 1]; }
        Vm = fr[12]; Vp = fr[13];
        bg = a.beta_gamma[c];
    } else {
        float mm = f0 < 0 ? 1.0f : 0.0f, pm = f0 > 0 ? 1.0f : 0.0f;
        for (int f = 0; f < 6; ++f) { am[f] = nominal[f] * mm; ap[f] = nominal[f] * pm; }
        Vm = vol * mm; Vp = vol * pm;
    }
    float cm[6], cp[6], sum_m = 0.0f, sum_p = 0.0f;
    for (int f = 0; f < 6; ++f) {
        cm[f] = am[f] * a.mu_m_faces[f * np + p] / dd[f];
        cp[f] = ap[f] * a.mu_p_faces[f * np + p] / dd[f];
    }
    sum_m = cm[0] + cm[1] + cm[2] + cm[3] + cm[4] + cm[5];
    sum_p = cp[0] + cp[1] + cp[2] + cp[3] + cp[4] + cp[5];
    if (a.coef26) {
        // coeffs_ of the point, the input of the learned preconditioner (discretization.py:337-339; evaluated
        // at box-boundary points too)
        for (int f = 0; f < 6; ++f) {
            a.coef26[(int64_t)(2 * f) * a.n_out + out] = cm[f];
            a.coef26[(int64_t)(2 * f + 1) * a.n_out + out] = cp[f];
            a.coef26[(int64_t)(14 + 2 * f) * a.n_out + out] = am[f];
            a.coef26[(int64_t)(15 + 2 * f) * a.n_out + out] = ap[f];
        }
        a.coef26[(int64_t)12 * a.n_out + out] = Vm;
        a.coef26[(int64_t)13 * a.n_out + out] = Vp;
    }
    if (bnd_early) return;
    float km = a.k_m[p], kp = a.k_p[p];
    float diag = kp * Vp + km * Vm + sum_m + sum_p;
    float rhs = a.f_m[p] * Vm + a.f_p[p] * Vp + bg;
    // unnormalised weights on u^-(slot), u^+(slot)
    float wm[7], wp[7];
    wm[0] = km * Vm + sum_m; wp[0] = kp * Vp + sum_p;
    for (int f = 0; f < 6; ++f) { wm[f + 1] = -cm[f]; wp[f + 1] = -cp[f]; }
    float inv = 1.0f / diag;
    bool ok = (diag != 0.0f) && isfinite(inv);
    float wU[7], wE[7];
    int32_t cE[7];
    bool irregular = false;
    uint8_t nlr = 0;
    float nlw = 0.0f, nl0 = 0.0f, nl1 = 0.0f;
    for (int s = 0; s < 7; ++s) {
        int8_t fs = a.flag[sid[s]];
        uint8_t sd = a.side[sid[s]];
        float u_w = 0.0f, e_w = 0.0f;
        cE[s] = -1;
        if (fs == -1) u_w = wm[s];
        else if (fs == 1) u_w = wp[s];
        else if (fs == 0) {
            // crossed: delta>0 -> (u^-,u^+) = (E,U) else (U,E)  (:484-485, :507-508)
            if (sd & 2) { e_w = wm[s]; u_w = wp[s]; }
            else { u_w = wm[s]; e_w = wp[s]; }
            cE[s] = a.cidx[sid[s]];
            irregular = true;
        }
        // the one-coefficient-per-face table assumes the row couples nodes of its own side only
        if (a.faces && fs != f0) irregular = true;
        wU[s] = ok ? u_w * inv : 0.0f;
        wE[s] = ok ? e_w * inv : 0.0f;
    }
    // nonlinear term N^-(u^-)V^- + N^+(u^+)V^+ (:369)
    if (ok) {
        if (f0 == 0) {
            uint8_t sd = a.side[sid[0]];
            if (sd & 2) { nl1 = Vp * inv; nlr = 1; nlw = Vm * inv; }   // u^+ = U, u^- = E
            else { nl0 = Vm * inv; nlr = 2; nlw = Vp * inv; }           // u^- = U, u^+ = E
        } else {
            nl0 = Vm * inv; nl1 = Vp * inv;
        }
    }
    // nan_to_num(rhs/diag) (:419): NaN -> 0, +-inf -> +-FLT_MAX
    float rn = ok ? rhs * inv : 0.0f;
    if (isnan(rn)) rn = 0.0f;
    else if (isinf(rn)) rn = rn > 0.0f ? 3.4028234664e38f : -3.4028234664e38f;
    if (a.faces) {
        // one coefficient per face, on the node's own side (0 for crossed cells: any row that touches one is in
        // the irregular list).  The -x face of the first plane of the slab belongs to the halo node.
        float own[6];
        for (int f = 0; f < 6; ++f) own[f] = f0 < 0 ? cm[f] : (f0 > 0 ? cp[f] : 0.0f);
        a.cface[out] = own[1];
        a.cface[a.n_out + out] = own[3];
        a.cface[2 * a.n_out + out] = own[5];
        if (i == 0) a.cface[out - a.out_stride[0]] = own[0];
        if (j == 0) a.cface[a.n_out + out - a.out_stride[1]] = own[2];
        if (k == 0) a.cface[2 * a.n_out + out - a.out_stride[2]] = own[4];
        if (bnd) {
            a.dinv[out] = -1.0f;                 // Dirichlet row: r = u - g (:389-393, :410-411)
            if (a.kv) a.kv[out] = 0.0f;
            a.rhs[out] = a.g_dir[p];
            if (a.nl) { a.nl[out] = 0.0f; a.nl[a.n_out + out] = 0.0f; }
            a.irr[out] = -1;
            return;
        }
        const bool dense = ok && !irregular;
        if (a.kv) a.kv[out] = dense ? (kp * Vp + km * Vm) : 0.0f;
        a.dinv[out] = dense ? inv : 0.0f;
        a.rhs[out] = dense ? rn : 0.0f;
        if (a.nl) { a.nl[out] = dense ? nl0 : 0.0f; a.nl[a.n_out + out] = dense ? nl1 : 0.0f; }
    } else {
        for (int s = 0; s < 7; ++s) a.w[s * a.n_out + out] = wU[s];
        a.rhs[out] = rn;
        if (a.nl) { a.nl[out] = nl0; a.nl[a.n_out + out] = nl1; }
    }
    int32_t slot = -1;
    if (irregular && ok) {
        unsigned long long q = atomicAdd((unsigned long long*)a.irr_count, 1ULL);
        if ((int64_t)q < a.irr_capacity) {
            slot = (int32_t)q;
            a.irr_point[q] = out;
            for (int s = 0; s < 7; ++s) { a.irr_wE[q * 7 + s] = wE[s]; a.irr_c[q * 7 + s] = cE[s]; }
            a.irr_nl[q] = nlr;
            a.irr_nlw[q] = nlw;
            if (a.faces) {
                for (int s = 0; s < 7; ++s) a.irr_wU[q * 7 + s] = wU[s];
                a.irr_rhs[q] = rn;
                // the U-part of the nonlinear term of an irregular row stays in the dense nl table
                if (a.nl) { a.nl[out] = nl0; a.nl[a.n_out + out] = nl1; }
            }
        }
    }
    a.irr[out] = slot;
}

}  // namespace nbm

using namespace nbm;

extern "C" {

const char* nbm_last_error(void) { return g_err; }
int nbm_version(void) { return 100; }

int nbm_ghost_layer_f32(const float* phi, const float* x, const float* y, const float* z, int nx, int ny, int nz,
                        float* phi_g, float* xg, float* yg, float* zg, nbm_stream_t stream) {
    NBM_REQUIRE(phi && x && y && z && phi_g && xg && yg && zg, "null pointer");
    NBM_REQUIRE(nx >= 3 && ny >= 3 && nz >= 3, "lvl grid needs >= 3 nodes per axis");
    cudaStream_t st = as_stream(stream);
    int gx = nx + 2, gy = ny + 2, gz = nz + 2;
    int rc = cuda_check(cudaMemsetAsync(phi_g, 0, sizeof(float) * (size_t)gx * gy * gz, st), "memset phi_g");
    if (rc) return rc;
    int64_t n = (int64_t)nx * ny * nz;
    ghost_fill_interior<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(phi, phi_g, nx, ny, nz);
    ghost_extrapolate<<<(unsigned)(((int64_t)ny * nz + 255) / 256), 256, 0, st>>>(phi_g, gx, gy, gz, 0);
    ghost_extrapolate<<<(unsigned)(((int64_t)gx * gz + 255) / 256), 256, 0, st>>>(phi_g, gx, gy, gz, 1);
    ghost_extrapolate<<<(unsigned)(((int64_t)gx * gy + 255) / 256), 256, 0, st>>>(phi_g, gx, gy, gz, 2);
    ghost_coords<<<(gx + 255) / 256, 256, 0, st>>>(x, xg, nx);
    ghost_coords<<<(gy + 255) / 256, 256, 0, st>>>(y, yg, ny);
    ghost_coords<<<(gz + 255) / 256, 256, 0, st>>>(z, zg, nz);
    NBM_LAUNCH_CHECK("ghost layer");
    return NBM_OK;
}

static int check_lvl(const nbm_lvl_t* lvl, bool sampled_ok = false) {
    NBM_REQUIRE(lvl, "null lvl");
    if (sampled_ok && !lvl->phi_g) return NBM_OK;   // sampled level set: the caller hands the values in
    NBM_REQUIRE(lvl->phi_g && lvl->xg && lvl->yg && lvl->zg, "null lvl grid");
    NBM_REQUIRE(lvl->gx >= 5 && lvl->gy >= 5 && lvl->gz >= 5, "ghosted lvl grid too small");
    if (lvl->interp != NBM_INTERP_TRILINEAR && lvl->interp != NBM_INTERP_QUADRATIC) {
        set_error("unknown interpolation kind %d", lvl->interp);
        return NBM_ERR_UNSUPPORTED;
    }
    return NBM_OK;
}

static int check_lat(const nbm_lattice_t* lat) {
    NBM_REQUIRE(lat && lat->xs && lat->ys && lat->zs, "null lattice");
    NBM_REQUIRE(lat->nx > 0 && lat->ny > 0 && lat->nz > 0, "empty lattice");
    NBM_REQUIRE(lat->n_shift >= 1 && lat->n_shift <= 7, "n_shift must be in 1..7");
    return NBM_OK;
}

int nbm_phi_interp_f32(const nbm_lvl_t* lvl, const float* pts, int64_t n, float* out, nbm_stream_t stream) {
    int rc = check_lvl(lvl);
    if (rc) return rc;
    NBM_REQUIRE(n >= 0, "negative n");
    if (n == 0) return NBM_OK;
    NBM_REQUIRE(pts && out, "null pointer");
    phi_interp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(*lvl, pts, n, out);
    NBM_LAUNCH_CHECK("phi_interp");
    return NBM_OK;
}

int nbm_classify_f32(const nbm_lvl_t* lvl, const nbm_lattice_t* lat, float dx, float dy, float dz, int8_t* flag,
                     uint8_t* side, nbm_stream_t stream) {
    int rc = check_lvl(lvl);
    if (rc) return rc;
    rc = check_lat(lat);
    if (rc) return rc;
    NBM_REQUIRE(flag && side, "null output");
    NBM_REQUIRE(dx > 0 && dy > 0 && dz > 0, "cell size must be positive");
    int64_t total = (int64_t)lat->nx * lat->ny * lat->nz * lat->n_shift;
    classify_kernel<<<(unsigned)((total + 127) / 128), 128, 0, as_stream(stream)>>>(*lvl, *lat, dx, dy, dz, flag, side,
                                                                                  total);
    NBM_LAUNCH_CHECK("classify");
    return NBM_OK;
}

int nbm_compact_crossed(const int8_t* flag, int64_t n, int64_t* idx_out, int64_t capacity, int32_t* cidx,
                        int64_t* count_dev, void* workspace, size_t* ws_bytes, nbm_stream_t stream) {
    NBM_REQUIRE(ws_bytes, "ws_bytes is null");
    NBM_REQUIRE(n > 0 && n < (int64_t)2147483647, "n out of range");
    cudaStream_t st = as_stream(stream);
    cub::CountingInputIterator<int64_t> counting(0);
    cub::TransformInputIterator<bool, IsZero, const int8_t*> sel(flag, IsZero());
    size_t need = 0;
    cudaError_t e = cub::DeviceSelect::Flagged(nullptr, need, counting, sel, idx_out, count_dev, (int)n, st);
    int rc = cuda_check(e, "cub size query");
    if (rc) return rc;
    if (!workspace) {
        *ws_bytes = need;
        return NBM_OK;
    }
    if (*ws_bytes < need) {
        set_error("workspace %zu < %zu", *ws_bytes, need);
        return NBM_ERR_WORKSPACE;
    }
    NBM_REQUIRE(flag && idx_out && cidx && count_dev, "null pointer");
    NBM_REQUIRE(capacity >= n, "idx_out must hold n entries (cub writes every selected item)");
    e = cub::DeviceSelect::Flagged(workspace, need, counting, sel, idx_out, count_dev, (int)n, st);
    rc = cuda_check(e, "cub select");
    if (rc) return rc;
    cidx_fill<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cidx, n);
    cidx_scatter<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(idx_out, count_dev, capacity, cidx);
    NBM_LAUNCH_CHECK("compact");
    return NBM_OK;
}

int nbm_cutcell_f32(const nbm_lvl_t* lvl, const nbm_lattice_t* lat, float dx, float dy, float dz, const int64_t* idx,
                    int64_t n_crossed, float* frac, float* tri, float* tri_area, nbm_stream_t stream) {
    int rc = check_lvl(lvl, lvl && lvl->corner_phi);
    if (rc) return rc;
    rc = check_lat(lat);
    if (rc) return rc;
    NBM_REQUIRE(n_crossed >= 0, "negative count");
    if (n_crossed == 0) return NBM_OK;
    NBM_REQUIRE(idx && frac && tri && tri_area, "null pointer");
    cutcell_kernel<<<(unsigned)((n_crossed + 63) / 64), 64, 0, as_stream(stream)>>>(*lvl, *lat, dx, dy, dz, idx,
                                                                                  n_crossed, frac, tri, tri_area);
    NBM_LAUNCH_CHECK("cutcell");
    return NBM_OK;
}

int nbm_regression_f32(const nbm_lvl_t* lvl, const nbm_lattice_t* lat, float dx, float dy, float dz,
                       const int64_t* idx, int64_t n_crossed, float* pos, float* proj, float* delta, float* Cm,
                       float* Cp, uint32_t* cube_side, nbm_stream_t stream) {
    int rc = check_lvl(lvl, lvl && lvl->cube_phi);
    if (rc) return rc;
    rc = check_lat(lat);
    if (rc) return rc;
    NBM_REQUIRE(n_crossed >= 0, "negative count");
    if (n_crossed == 0) return NBM_OK;
    NBM_REQUIRE(idx && pos && proj && delta && Cm && Cp && cube_side, "null pointer");
    regression_kernel<<<(unsigned)((n_crossed + 63) / 64), 64, 0, as_stream(stream)>>>(
        *lvl, *lat, dx, dy, dz, idx, n_crossed, pos, proj, delta, Cm, Cp, cube_side);
    NBM_LAUNCH_CHECK("regression");
    return NBM_OK;
}

int nbm_site_weights_f32(int64_t n_crossed, const float* delta, const float* Cm, const float* Cp,
                         const float* mu_m_s, const float* mu_p_s, const float* alpha_proj, const float* beta_proj,
                         const float* mu_m_proj, const float* mu_p_proj, float* B, nbm_stream_t stream) {
    NBM_REQUIRE(n_crossed >= 0, "negative count");
    if (n_crossed == 0) return NBM_OK;
    NBM_REQUIRE(delta && Cm && Cp && mu_m_s && mu_p_s && alpha_proj && beta_proj && mu_m_proj && mu_p_proj && B,
                "null pointer");
    site_weights_kernel<<<(unsigned)((n_crossed + 63) / 64), 64, 0, as_stream(stream)>>>(
        n_crossed, delta, Cm, Cp, mu_m_s, mu_p_s, alpha_proj, beta_proj, mu_m_proj, mu_p_proj, B);
    NBM_LAUNCH_CHECK("site_weights");
    return NBM_OK;
}

int nbm_assemble_f32(const nbm_assemble_t* a, nbm_stream_t stream) {
    NBM_REQUIRE(a, "null plan");
    int rc = check_lat(&a->pts);
    if (rc) return rc;
    NBM_REQUIRE(a->flag && a->side && a->cidx, "null site tables");
    NBM_REQUIRE(a->mu_m_faces && a->mu_p_faces && a->k_m && a->k_p && a->f_m && a->f_p && a->g_dir,
                "null coefficient samples");
    NBM_REQUIRE(a->rhs && a->irr, "null outputs");
    if (a->faces) {
        NBM_REQUIRE(a->shared, "faces mode needs the shared (lattice) layout");
        NBM_REQUIRE(a->cface && a->dinv && a->irr_wU && a->irr_rhs, "null face-table outputs");
    } else {
        NBM_REQUIRE(a->w, "null outputs");
    }
    NBM_REQUIRE(a->irr_count && a->irr_point && a->irr_wE && a->irr_c && a->irr_nl && a->irr_nlw,
                "null irregular-row buffers");
    NBM_REQUIRE(a->dx > 0 && a->dy > 0 && a->dz > 0, "cell size must be positive");
    int64_t np = (int64_t)a->pts.nx * a->pts.ny * a->pts.nz;
    assemble_kernel<<<(unsigned)((np + 127) / 128), 128, 0, as_stream(stream)>>>(*a);
    NBM_LAUNCH_CHECK("assemble");
    return NBM_OK;
}

}  // extern "C"
