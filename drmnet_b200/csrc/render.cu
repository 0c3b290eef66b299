// K1: reflectance-map forward render as a tiled gather-reduce over the environment map (sm_100a).
//
// Replaces MitsubaRefMapRenderer.rendering (reference utils/mitsuba3_utils.py:411-430 -> :365-409 -> :217-246): the
// Mitsuba scene {unit sphere, `principled` BSDF, lat-long envmap emitter, normal-indexed orthographic sensor, `direct`
// integrator, box filter} evaluated as its deterministic limit
//
//   out[k, i, j, c] = sum_{a,b < S} w_a w_b  sum_texels  f_c(d_t; v_k, n(theta_i,a, phi_j,b); z_k) E[t, c] dOmega_t
//
// with S x S Gauss-Legendre sub-normals per refmap cell (the box pixel filter) and the texel-centre quadrature of the
// emitter.  One CTA owns a tile of refmap cells of one render and streams 32x32-texel envmap tiles through shared
// memory with TMA (cp.async.bulk.tensor, double buffered, mbarrier completion); a cooperative transform turns each raw
// tile into per-texel records (half vector, Fresnel-weighted radiance * solid angle, retro-reflection factor) that are
// pixel independent, then every thread gathers records for its 4 sub-normal slots from shared memory (LDS.128).
//
// Controlled reductions of the pair count keep the result within ~6e-5 of the full sum (tests):
//   * footprint levels: a texel tile far (in half-vector space) from the CTA's normals is integrated over the cell
//     with a coarser Gauss-Legendre lattice (16x16 -> 8x8 -> 4x4 -> 2x2 -> 1x1); the slots freed by the coarser
//     lattice split the tile's texels among themselves, so no thread idles;
//   * for the 8x8 and 16x16 footprints the near field is split off per (cell, texel) and evaluated by
//     render_near_kernel (one CTA per cell, window scan around the mirror direction), the tile kernel keeps the 1x1 rest;
//   * the diffuse lobe, which varies on the scale of a radian, is gathered from a 4x4-texel energy-centroid coarsening
//     of the envmap (built per call by env_coarsen_kernel) when those cells are small enough;
//   * a rough specular lobe (alpha >= cell size / 0.018) is gathered from the 4x4 or the 2x2 coarsening.
// The kernel is bound by the FP32/MUFU pipes (about 22 instructions per (slot, texel) pair for the specular lobe, 14
// for the diffuse one), not by HBM: each envmap byte is reused by every slot of the render out of L2.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

// This file is the round-1 tile kernel kept as the SINGLE-LEVEL validation path (drm_render_refmaps_flat): every
// switch of its hierarchy is off, no environment variable is read.
static const char* flat_switch(const char* name) {
    return (strstr(name, "SCALE") || strstr(name, "SLOTS")) ? nullptr : "0";
}
#define getenv(name) flat_switch(name)

namespace drm {

static constexpr int GATHER_THREADS = 256;
static constexpr int SUBS_PER_THREAD = 4;
static constexpr int SLOTS = GATHER_THREADS * SUBS_PER_THREAD;  // sub-normal slots per CTA
static constexpr int TT = 32;                                   // largest tile edge in texels (or coarse cells)
static constexpr int TILE_TEXELS = TT * TT;
static constexpr int REC_FLOATS = 12;
static constexpr int MAX_LEVELS = 5;   // footprint lattices 1,2,4,8,16 per axis
static constexpr int MAX_LIST = 2048;  // tiles one CTA can schedule (the plan splits larger maps)
static constexpr int COARSE = 4;       // coarsening factor of the energy-centroid map (diffuse lobe, very rough lobes)
static constexpr int COARSE2 = 2;      // finer coarsening for moderately rough specular lobes
static constexpr int COARSE_FLOATS = 6;  // centroid direction + radiance * solid angle per coarse cell
static constexpr int FAR_EDGE = 16;     // cells per side of the far launch's blocks
static constexpr int FARC_TT = 16;      // tile edge (coarse cells) of the far-field launches on the 2x2 coarse map

// A launch pair splits the (cell, tile) plane: the far launch takes, with large blocks of cells and the 1x1 lattice,
// every tile that is far (1x1-accurate) from the whole block; the near launch takes the rest with small blocks.
enum : int { PART_ALL = 0, PART_FAR = 1, PART_NEAR = 2 };

// which launch serves which part of a render
enum : int { ROUTE_SPEC_RAW = 1, ROUTE_BOTH_RAW = 2, ROUTE_DIFF_COARSE = 4, ROUTE_BOTH_COARSE = 8, ROUTE_BOTH_COARSE2 = 16 };

struct RenderConst {  // per render
    float vhat[3], left[3], upp[3];  // camera frame of look_at(v, 0, +Y); `left` carries the flip sign
    float m, rough, alpha2, inv_a2m1, one_m_a2, eta;
    float base[3], cdiff[3];
    float thr[MAX_LEVELS];  // half-vector-space distance beyond which footprint level k is accurate enough
    float tk2[MAX_LEVELS];  // the same for one cell: squared chord |n_centre - h|^2 thresholds (cell radius included)
    float near_reach;       // angle of the chord tk2[0]: nothing closer may leave the raw map in near mode
    int env, route;
};

struct GatherArgs {
    const float* src;  // raw envmaps [B,He,We,3] or coarse maps [B,Hm,Wm,6]
    const RenderConst* rc;
    const float *sin_t, *cos_t, *sin_p, *cos_p;
    float* slab;  // [splits][N][res*res][3] partial sums of this launch
    int B, He, We, N, res, S;
    int Hm, Wm;   // rows / columns of the map this launch reads (He,We or the coarse dims)
    int tile_w, tile_h, tiles_x, tiles_y;
    int ttiles_x, ttiles_y, splits;
    int use_tma, cull, route_mask;
    int tt;                    // tile edge in texels of this launch's map (8, 16 or 32)
    int G;                     // sub-normal slots per refmap cell (a multiple of S*S)
    int part;                  // PART_ALL, or the far / near half of a launch pair (see far_tile)
    int far_edge;              // edge, in cells, of the blocks the far launch works on
    int far_mode;              // 0: off; 1: raw-map launch skips tiles beyond dfar; 2: coarse-map launch keeps only those
    float dfar;                // distance (block of cells to raw tile, half-vector space) beyond which the 2x2 map serves
    float dfar_hi;             // far_mode 2: ... and keeps only tiles closer than this (0: no upper end; the next coarser map)
    int far_factor;            // far_mode 2: texels per cell edge of this launch's map (2 or 4)
    int dfar_near;             // near mode: ... and never closer than the render's near_reach (render_near_kernel's pairs)
    int raw_tt, raw_ttiles_x, raw_Hm, raw_Wm;  // far_mode 2: geometry of the raw map's tiles (the unit of the decision)
    float raw_dth, raw_dph;
    int nlev;                  // number of footprint levels used by this launch
    int lev_S[MAX_LEVELS];     // lattice size per axis of level k (ascending; the last one is S)
    int lev_tidx[MAX_LEVELS];  // log2(lev_S[k]): index into RenderConst::thr
    float domega_k, cell, dth_cell, dph_cell;  // dth/dph: angular size of one map cell
    float gl_x[MAX_LEVELS][16], gl_w[MAX_LEVELS][16];
    int fine_S;                // the render's own lattice (>= every lev_S) and its nodes: view_term_avg
    float fine_x[16], fine_w[16];
};

__global__ void render_tables_kernel(float* sin_t, float* cos_t, float* sin_p, float* cos_p, int He, int We) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < He) {
        double t = (i + 0.5) * (M_PI / He);
        sin_t[i] = (float)sin(t);
        cos_t[i] = (float)cos(t);
    }
    if (i < We) {
        double p = (i + 0.5) * (2.0 * M_PI / We);
        sin_p[i] = (float)sin(p);
        cos_p[i] = (float)cos(p);
    }
}

// K3: energy-centroid coarsenings, 2x2 and 4x4 texels per cell, in one pass over the envmaps this call uses.
// Cell = {unit centroid direction (luminance * solid-angle weighted), sum of radiance * solid angle per channel}: placing
// the cell's energy at its centroid cancels the first-order error of evaluating a smooth lobe once per cell.
// One thread per 4x4 cell: it writes its four 2x2 cells and their union.
__global__ void env_mark_used_kernel(const RenderConst* __restrict__ rc, int N, int* __restrict__ used) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) used[rc[k].env] = 1;
}

__global__ void env_coarsen_kernel(const float* __restrict__ env, const float* __restrict__ sin_t,
                                   const float* __restrict__ cos_t, const float* __restrict__ sin_p,
                                   const float* __restrict__ cos_p, const int* __restrict__ used, int B, int He, int We,
                                   int Hc, int Wc, int Hc2, int Wc2, float domega_k, float* __restrict__ coarse4,
                                   float* __restrict__ coarse2) {
    const long cell = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (cell >= (long)B * Hc * Wc) return;
    const int C = (int)(cell % Wc), R = (int)((cell / Wc) % Hc), b = (int)(cell / ((long)Wc * Hc));
    if (!used[b]) return;
    const float* src = env + (size_t)b * He * We * 3;
    float M0 = 0.f, M1 = 0.f, M2 = 0.f, CX = 0.f, CY = 0.f, CZ = 0.f, GX = 0.f, GY = 0.f, GZ = 0.f;
    for (int sr = 0; sr < 2; ++sr)
        for (int sc = 0; sc < 2; ++sc) {
            const int R2 = R * 2 + sr, C2 = C * 2 + sc;
            if (R2 >= Hc2 || C2 >= Wc2) continue;
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, cx = 0.f, cy = 0.f, cz = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
            for (int dr = 0; dr < 2; ++dr) {
                const int r = R2 * 2 + dr;
                if (r >= He) break;
                const float st = sin_t[r], ct = cos_t[r], dom = domega_k * st;
                for (int dc = 0; dc < 2; ++dc) {
                    const int c = C2 * 2 + dc;
                    if (c >= We) break;
                    const float* e = src + ((size_t)r * We + c) * 3;
                    const float er = e[0] * dom, eg = e[1] * dom, eb = e[2] * dom;
                    const float dx = st * sin_p[c], dy = ct, dz = -st * cos_p[c];
                    const float w = er + eg + eb;
                    m0 += er; m1 += eg; m2 += eb;
                    cx += w * dx; cy += w * dy; cz += w * dz;
                    gx += dom * dx; gy += dom * dy; gz += dom * dz;  // geometric centre, used when the cell is black
                }
            }
            M0 += m0; M1 += m1; M2 += m2;
            CX += cx; CY += cy; CZ += cz;
            GX += gx; GY += gy; GZ += gz;
            float n2 = cx * cx + cy * cy + cz * cz;
            if (!(n2 > 1e-30f)) { cx = gx; cy = gy; cz = gz; n2 = cx * cx + cy * cy + cz * cz; }
            const float inv = rsqrtf(fmaxf(n2, 1e-38f));
            float* o = coarse2 + (((size_t)b * Hc2 + R2) * Wc2 + C2) * COARSE_FLOATS;
            o[0] = cx * inv; o[1] = cy * inv; o[2] = cz * inv;
            o[3] = m0; o[4] = m1; o[5] = m2;
        }
    float n2 = CX * CX + CY * CY + CZ * CZ;
    if (!(n2 > 1e-30f)) { CX = GX; CY = GY; CZ = GZ; n2 = CX * CX + CY * CY + CZ * CZ; }
    const float inv = rsqrtf(fmaxf(n2, 1e-38f));
    float* o = coarse4 + (size_t)cell * COARSE_FLOATS;
    o[0] = CX * inv; o[1] = CY * inv; o[2] = CZ * inv;
    o[3] = M0; o[4] = M1; o[5] = M2;
}

// 16x16-texel cells from the 4x4 ones (only the lattice correction of the diffuse lobe reads them)
__global__ void env_coarsen16_kernel(const float* __restrict__ coarse4, const int* __restrict__ used, int B, int Hc, int Wc,
                                     int Hc16, int Wc16, float* __restrict__ coarse16) {
    const long cell = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (cell >= (long)B * Hc16 * Wc16) return;
    const int C = (int)(cell % Wc16), R = (int)((cell / Wc16) % Hc16), b = (int)(cell / ((long)Wc16 * Hc16));
    if (!used[b]) return;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, cx = 0.f, cy = 0.f, cz = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    for (int dr = 0; dr < 4; ++dr) {
        const int r = R * 4 + dr;
        if (r >= Hc) break;
        for (int dc = 0; dc < 4; ++dc) {
            const int c = C * 4 + dc;
            if (c >= Wc) break;
            const float* e = coarse4 + (((size_t)b * Hc + r) * Wc + c) * COARSE_FLOATS;
            const float w = e[3] + e[4] + e[5];
            m0 += e[3]; m1 += e[4]; m2 += e[5];
            cx += w * e[0]; cy += w * e[1]; cz += w * e[2];
            gx += e[0]; gy += e[1]; gz += e[2];
        }
    }
    float n2 = cx * cx + cy * cy + cz * cz;
    if (!(n2 > 1e-30f)) { cx = gx; cy = gy; cz = gz; n2 = cx * cx + cy * cy + cz * cz; }
    const float inv = rsqrtf(fmaxf(n2, 1e-38f));
    float* o = coarse16 + (size_t)cell * COARSE_FLOATS;
    o[0] = cx * inv; o[1] = cy * inv; o[2] = cz * inv;
    o[3] = m0; o[4] = m1; o[5] = m2;
}

// clip z to [0,1] (mitsuba3_utils.py:239,242), derive the BSDF constants, the camera frame (:235-236), the footprint
// level thresholds and the routing of the render's terms to the launches
__global__ void render_setup_kernel(const float* __restrict__ z6, const float* __restrict__ view3,
                                    const uint8_t* __restrict__ flip, const int32_t* __restrict__ env_index, int N,
                                    int B, float alpha_min, float cell, float level_scale, float near_scale,
                                    float coarse_h, float coarse2_h, int coarse_diffuse_ok, int unify_coarse2,
                                    RenderConst* __restrict__ rc) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    RenderConst c;
    float z[6];
    for (int i = 0; i < 6; ++i) z[i] = fminf(fmaxf(z6[6 * k + i], 0.f), 1.f);
    c.m = z[0];
    c.base[0] = z[1]; c.base[1] = z[2]; c.base[2] = z[3];
    c.rough = z[4];
    float alpha = fmaxf(z[4] * z[4], alpha_min);
    c.alpha2 = alpha * alpha;
    c.inv_a2m1 = 1.f / c.alpha2 - 1.f;
    c.one_m_a2 = 1.f - c.alpha2;
    c.eta = 2.f / (1.f - sqrtf(0.08f * z[5])) - 1.f;
    // Distance (angle between the cell's normal and the half vector, radians) beyond which a coarser footprint
    // lattice integrates the GGX tail ~ (alpha^2 + d^2)^-2 over the cell accurately enough: the S-point
    // Gauss-Legendre error ~ K_S (cell/d)^(2S) weighted by the tail mass (alpha/d)^2 is held near 1e-5.
    {
        const float ca = cell * alpha;
        c.thr[0] = 21.0f * sqrtf(ca);                                      // 1 x 1
        c.thr[1] = 7.5f * powf(cell, 2.f / 3.f) * powf(alpha, 1.f / 3.f);  // 2 x 2
        c.thr[2] = 2.4f * powf(cell, 0.8f) * powf(alpha, 0.2f);            // 4 x 4
        c.thr[3] = 1.2f * powf(cell, 8.f / 9.f) * powf(alpha, 1.f / 9.f);  // 8 x 8
        c.thr[4] = 0.f;                                                    // 16 x 16
        for (int i = 0; i < MAX_LEVELS - 1; ++i) c.thr[i] = level_scale * fmaxf(c.thr[i], 6.f * alpha);  // beyond the core
        for (int i = 0; i < MAX_LEVELS; ++i) {
            // per-cell thresholds are applied exactly (no tile / block slack), hence their own, larger scale
            const float t = c.thr[i] * (near_scale / level_scale) + 0.75f * cell;  // nodes lie within 0.75 cell of the centre
            c.tk2[i] = i < MAX_LEVELS - 1 ? t * t : 0.f;
        }
        c.near_reach = 2.f * asinf(fminf(1.f, 0.5f * sqrtf(c.tk2[0]))) + 2e-3f;
    }
    for (int i = 0; i < 3; ++i) c.cdiff[i] = (1.f - c.m) * c.base[i] * (float)M_1_PI;
    const bool has_diffuse = c.cdiff[0] > 0.f || c.cdiff[1] > 0.f || c.cdiff[2] > 0.f;
    // routing: the error of one evaluation per coarse cell of size h is ~ 0.03 (h/alpha)^2 for the GGX lobe on smooth
    // skies and up to ~ 6e-6 / alpha^2 for a compact sun inside a cell (its extent is lost), ~ 0.18 h^2 for the diffuse
    // lobe (oracle and scripts/coarse_probe.py studies); h <= 0.018 alpha holds the worst map near 5e-5
    const bool spec_coarse = coarse_h > 0.f && coarse_h <= 0.018f * alpha;
    const bool spec_coarse2 = coarse2_h > 0.f && coarse2_h <= 0.018f * alpha;
    if (spec_coarse) c.route = ROUTE_BOTH_COARSE;
    // (with the distance-switched far launches the 2x2 rule is their d = 0 case: such a render takes the raw-map
    // route, whose raw launch then finds no tile, with the 4x4 map beyond its switch distance and for the diffuse lobe)
    else if (spec_coarse2 && !unify_coarse2) c.route = ROUTE_BOTH_COARSE2;
    else if (!has_diffuse) c.route = ROUTE_SPEC_RAW;
    else c.route = coarse_diffuse_ok ? (ROUTE_SPEC_RAW | ROUTE_DIFF_COARSE) : ROUTE_BOTH_RAW;
    float vx = view3[3 * k], vy = view3[3 * k + 1], vz = view3[3 * k + 2];
    float inv = rsqrtf(vx * vx + vy * vy + vz * vz);
    vx *= inv; vy *= inv; vz *= inv;
    c.vhat[0] = vx; c.vhat[1] = vy; c.vhat[2] = vz;
    float fx = -vx, fy = -vy, fz = -vz;  // forward
    float lx = fz, ly = 0.f, lz = -fx;   // up x forward, up = (0,1,0)
    float linv = rsqrtf(lx * lx + lz * lz);
    lx *= linv; lz *= linv;
    c.upp[0] = fy * lz - fz * ly;  // forward x left
    c.upp[1] = fz * lx - fx * lz;
    c.upp[2] = fx * ly - fy * lx;
    float sgn = (flip && flip[k]) ? -1.f : 1.f;
    c.left[0] = sgn * lx; c.left[1] = sgn * ly; c.left[2] = sgn * lz;
    int e = env_index ? env_index[k] : k;
    c.env = min(max(e, 0), B - 1);
    rc[k] = c;
}

// The view-side factor of the specular lobe, G1(n.v) / (4 n.v) D's normalisation included, depends on the normal only
// and varies on the scale alpha at the limb (n.v -> 0), however far the texel is.  A node of a coarser lattice stands
// for the m x m nodes of the render's own S x S lattice that lie in its sub-region (Gauss-Legendre nodes of order 2n
// interlace the weight intervals of order n: m consecutive nodes per axis); it carries their weighted mean of this
// factor, so the footprint levels approximate only the texel-dependent part.  m = 1 returns the factor itself.
__device__ __forceinline__ float view_term(const RenderConst& rc, float lz) {
    const float g1 = lz + fast_sqrt(lz * lz * rc.one_m_a2 + rc.alpha2);
    return lz > 0.f ? fast_rcp(3.14159265358979f * rc.alpha2 * g1) : 0.f;
}

// The node also moves by the shift of the region's centroid that this weighting causes (weighted minus unweighted
// centroid of the fine nodes, in lattice coordinates): the first-order cross term between the view factor and the
// texel-dependent part then cancels, and where the factor is flat the Gauss-Legendre node stays where it is.
// fx(n), fw(n): node / weight n of the finest lattice (functors, so kernel parameters stay in the constant bank).
// Returns {mean factor, shift along theta, shift along phi}.
template <class FX, class FW>
__device__ __forceinline__ float3 view_term_avg(const RenderConst& rc, float cell, int i, int j, int a, int b, int m,
                                                FX fx, FW fw) {
    float num = 0.f, den = 0.f, va = 0.f, vb = 0.f, ua = 0.f, ub = 0.f;
    float sps[16];
    for (int ib = 0; ib < m; ++ib) sps[ib] = sinf(((float)j + 0.5f + 0.5f * fx(b * m + ib)) * cell);
    for (int ia = 0; ia < m; ++ia) {
        const float xa = fx(a * m + ia), wa = fw(a * m + ia);
        const float st = sinf(((float)i + 0.5f + 0.5f * xa) * cell);
        for (int ib = 0; ib < m; ++ib) {
            const float xb = fx(b * m + ib);
            const float sp = sps[ib];
            const float w = wa * fw(b * m + ib);
            const float wv = w * view_term(rc, st * sp);
            num += wv; den += w;
            va += wv * xa; vb += wv * xb;
            ua += w * xa; ub += w * xb;
        }
    }
    if (!(num > 0.f)) return make_float3(0.f, 0.f, 0.f);
    return make_float3(num / den, va / num - ua / den, vb / num - ub / den);
}

// Distance (block of cells to raw tile, half-vector space) from which a coarsening whose base distance is D serves a
// render: one evaluation per cell of half size h is off by ~ (h^2/24) 20 / (d^2 + 0.68 alpha^2), the tail's curvature
// with the lobe's flat core folded in (at d = 0 this is the whole-map rule h <= 0.018 alpha of the routing).
__device__ __forceinline__ float far_switch(const GatherArgs& g, const RenderConst& rc, float D) {
    const float d = sqrtf(fmaxf(D * D - 0.676f * rc.alpha2, 0.f));
    return g.dfar_near ? fmaxf(d, rc.near_reach) : d;
}

__device__ __forceinline__ bool far_in_range(const GatherArgs& g, const RenderConst& rc, float dr) {
    return dr >= far_switch(g, rc, g.dfar) && (g.dfar_hi <= 0.f || dr < far_switch(g, rc, g.dfar_hi));
}

__device__ __forceinline__ float fresnel_dielectric(float cos_i, float eta) {
    float eta_ti = 1.f / eta;
    float ct2 = 1.f - eta_ti * eta_ti * (1.f - cos_i * cos_i);
    if (ct2 <= 0.f) return 1.f;
    float ct = sqrtf(ct2);
    float a_s = (cos_i - eta * ct) / (cos_i + eta * ct);
    float a_p = (ct - eta * cos_i) / (ct + eta * cos_i);
    return 0.5f * (a_s * a_s + a_p * a_p);
}

// Cone (axis, angular radius) of the normals of the cells [i0,i1) x [j0,j1), cell corners included.
__device__ __forceinline__ void cell_block_cone(const GatherArgs& g, const RenderConst& rc, int i0, int i1, int j0,
                                                int j1, float& ax, float& ay, float& az, float& beta) {
    const float thc = 0.5f * (i0 + i1) * g.cell, phc = 0.5f * (j0 + j1) * g.cell;
    float st, ct, sp, cp;
    sincosf(thc, &st, &ct);
    sincosf(phc, &sp, &cp);
    const float lx = st * cp, lz = st * sp;
    ax = lx * rc.left[0] + ct * rc.upp[0] + lz * rc.vhat[0];
    ay = lx * rc.left[1] + ct * rc.upp[1] + lz * rc.vhat[1];
    az = lx * rc.left[2] + ct * rc.upp[2] + lz * rc.vhat[2];
    // angular radius: centre-to-corner distance d obeys d^2 <= dth^2 + (dph sin_max)^2 (5 % margin)
    const float dth = 0.5f * (i1 - i0) * g.cell, dph = 0.5f * (j1 - j0) * g.cell * fminf(1.f, st + dth);
    beta = (dth < 0.1f && dph < 0.1f) ? 1.05f * sqrtf(dth * dth + dph * dph) : dth + dph;  // small blocks: flat-space bound
}

// Distance, in half-vector space, between a cone of normals and the half vectors h = normalize(v + d) of one map tile;
// -1 when no normal of the cone sees any direction of the tile (n.d <= 0 everywhere), -0.5 when the d -> h map is
// singular over the tile (d ~ -v) and no bound is available.
__device__ __forceinline__ float tile_distance_geom(int cull, int tt, int ttiles_x, int Hm, int Wm, float dth_cell,
                                                    float dph_cell, const float* __restrict__ vhat, int tile, float ax,
                                                    float ay, float az, float beta) {
    const int ty = tile / ttiles_x, tx = tile - ty * ttiles_x;
    const int r0 = ty * tt, r1 = min(r0 + tt, Hm), c0 = tx * tt, c1 = min(c0 + tt, Wm);
    const float dth = 0.5f * (r1 - r0) * dth_cell, dph = 0.5f * (c1 - c0) * dph_cell;
    const float thc = fminf(0.5f * (r0 + r1) * dth_cell, 3.14159265f), phc = 0.5f * (c0 + c1) * dph_cell;
    float st, ct, sp, cp;
    sincosf(thc, &st, &ct);
    sincosf(phc, &sp, &cp);
    // angular radius of the tile: centre-to-corner distance d obeys d^2 <= dth^2 + (dph sin_max)^2 (5 % margin)
    const float dps = dph * fminf(1.f, st + dth);
    const float gamma = (dth < 0.1f && dps < 0.1f) ? 1.05f * sqrtf(dth * dth + dps * dps) : dth + dps;
    const float dx = st * sp, dy = ct, dz = -st * cp;
    if (cull) {
        const float spread = beta + gamma + 0.01f;
        if (spread < 1.5607963f && ax * dx + ay * dy + az * dz <= -sinf(spread)) return -1.f;
    }
    const float hx = vhat[0] + dx, hy = vhat[1] + dy, hz = vhat[2] + dz;
    const float len = sqrtf(hx * hx + hy * hy + hz * hz);
    if (len - gamma < 0.05f) return -0.5f;
    const float gamma_h = gamma / (len - gamma);  // |dh| <= |dd| / |v + d|
    const float cosang = fminf(fmaxf((ax * hx + ay * hy + az * hz) / len, -1.f), 1.f);
    return fmaxf(acosf(cosang) - beta - gamma_h, 0.f);
}

__device__ __forceinline__ float tile_distance(const GatherArgs& g, const float* __restrict__ vhat, int tile, float ax,
                                               float ay, float az, float beta) {
    return tile_distance_geom(g.cull, g.tt, g.ttiles_x, g.Hm, g.Wm, g.dth_cell, g.dph_cell, vhat, tile, ax, ay, az, beta);
}

// Classify one map tile for the CTA whose cells are [pi0,pi1) x [pj0,pj1):
//   0      skipped (invisible, or it belongs to the other launch of a far/near pair)
//   1+k    footprint level k (0 = 1x1 lattice ... nlev-1 = full S x S lattice)
__device__ __forceinline__ int classify_tile(const GatherArgs& g, const RenderConst& rcs, const float* __restrict__ vhat,
                                             const float* __restrict__ thr, int tile, int pi0, int pi1, int pj0, int pj1,
                                             float ax, float ay, float az, float beta) {
    if (g.part != PART_ALL) {
        // the far block enclosing (or equal to) my cells decides, identically in both launches of the pair
        const int e = g.far_edge;
        const int fi0 = (pi0 / e) * e, fj0 = (pj0 / e) * e;
        float fx, fy, fz, fb;
        cell_block_cone(g, rcs, fi0, min(fi0 + e, g.res), fj0, min(fj0 + e, g.res), fx, fy, fz, fb);
        const float fd = tile_distance(g, vhat, tile, fx, fy, fz, fb);
        const bool is_far = fd >= thr[0];
        if (g.part == PART_FAR) return (fd >= 0.f || fd == -0.5f) && is_far ? 1 : 0;
        if (is_far || fd == -1.f) return 0;  // far: the other launch; invisible from the big block: from mine too
    }
    const float dist = tile_distance(g, vhat, tile, ax, ay, az, beta);
    if (dist == -1.f) return 0;
    if (g.far_mode == 1 && dist >= far_switch(g, rcs, g.dfar)) return 0;  // served from the 2x2 coarse map by the far_mode 2 launch
    if (g.far_mode == 2) {
        // a coarse tile none of whose raw tiles is far for this block has nothing to contribute
        const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
        const int rt0 = (ty * g.tt * g.far_factor) / g.raw_tt, ct0 = (tx * g.tt * g.far_factor) / g.raw_tt;
        const int nrt = (g.tt * g.far_factor + g.raw_tt - 1) / g.raw_tt;
        const int raw_ttiles_y = (g.raw_Hm + g.raw_tt - 1) / g.raw_tt;
        bool any_far = false;
        for (int a = 0; a < nrt && !any_far; ++a)
            for (int b = 0; b < nrt && !any_far; ++b) {
                const int rt = rt0 + a, ct = ct0 + b;
                if (rt >= raw_ttiles_y || ct >= g.raw_ttiles_x) continue;
                any_far = far_in_range(g, rcs, tile_distance_geom(g.cull, g.raw_tt, g.raw_ttiles_x, g.raw_Hm, g.raw_Wm, g.raw_dth,
                                                                  g.raw_dph, vhat, rt * g.raw_ttiles_x + ct, ax, ay, az, beta));
            }
        if (!any_far) return 0;
    }
    if (g.nlev == 1) return 1;
    if (dist == -0.5f) return g.nlev;  // no bound: stay on the finest lattice
    for (int k = 0; k < g.nlev - 1; ++k)
        if (dist >= thr[g.lev_tidx[k]]) return 1 + k;
    return g.nlev;
}

// TERMS: 1 = specular lobe, 2 = diffuse lobe, 3 = both.  COARSE_SRC: the map is an energy-centroid coarsening.
// NEAR_EXCL: 1x1 lattice only, and (cell, texel) pairs closer than the 1x1 threshold are left to render_near_kernel.
template <int TERMS, bool COARSE_SRC, bool NEAR_EXCL>
__global__ void __launch_bounds__(GATHER_THREADS, 2)
render_gather_kernel(const __grid_constant__ CUtensorMap tmap, const GatherArgs g) {
    constexpr int CELL_FLOATS = COARSE_SRC ? COARSE_FLOATS : 3;
    constexpr int RAW_FLOATS = TILE_TEXELS * CELL_FLOATS;  // stage size for the largest tile
    constexpr bool SPEC = (TERMS & 1) != 0, DIFF = (TERMS & 2) != 0;
    const int tt = g.tt, tile_texels = tt * tt, row_floats = tt * CELL_FLOATS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* raw0 = reinterpret_cast<float*>(smem_raw);
    float4* rec = reinterpret_cast<float4*>(smem_raw + 2 * RAW_FLOATS * sizeof(float));
    unsigned char* tail = smem_raw + 2 * RAW_FLOATS * sizeof(float) + TILE_TEXELS * REC_FLOATS * sizeof(float);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);                        // 2 mbarriers
    int* scan_ws = reinterpret_cast<int*>(tail + 16);                            // 8 warp sums + level ends
    uint16_t* list = reinterpret_cast<uint16_t*>(tail + 16 + 64);                // [MAX_LIST] tiles ordered by level
    uint8_t* lvl = reinterpret_cast<uint8_t*>(tail + 16 + 64 + MAX_LIST * 2);    // [MAX_LIST]

    const int tid = threadIdx.x;
    const int k = blockIdx.y;
    const RenderConst rc = g.rc[k];
    if (!(rc.route & g.route_mask)) return;  // another launch serves this render (uniform per CTA)

    const int ptile = blockIdx.x;
    const int pty = ptile / g.tiles_x, ptx = ptile - pty * g.tiles_x;
    const int pi0 = pty * g.tile_h, pj0 = ptx * g.tile_w;
    const int G = g.G;  // slots per cell
    const int npix = g.tile_w * g.tile_h;

    // ---- cone of this CTA's normals (cell corners included) ---------------------------------------------------------
    const int pi1 = min(pi0 + g.tile_h, g.res), pj1 = min(pj0 + g.tile_w, g.res);
    float ax, ay, az, beta;
    cell_block_cone(g, rc, pi0, pi1, pj0, pj1, ax, ay, az, beta);

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        if (g.use_tma) tma_prefetch_desc(&tmap);
    }

    // ---- schedule: classify my tiles, then order them by footprint level (coarse first, tile order inside) ----------
    // texel splits interleave the tiles (tile = z + e * splits) so that the few tiles near the lobe, which run on
    // the fine lattices, spread over all splits
    const int ntiles = g.ttiles_x * g.ttiles_y;
    const int tbeg = blockIdx.z, tstep = g.splits;
    const int nmine = (ntiles - tbeg + tstep - 1) / tstep;
    for (int e = tid; e < nmine; e += GATHER_THREADS)
        lvl[e] = (uint8_t)classify_tile(g, rc, g.rc[k].vhat, g.rc[k].thr, tbeg + e * tstep, pi0, pi1, pj0, pj1, ax, ay, az,
                                        beta);
    __syncthreads();
    int nlist = 0;
    {
        constexpr int PER_T = MAX_LIST / GATHER_THREADS;
        const int lane = tid & 31, w = tid >> 5;
        for (int L = 1; L <= g.nlev; ++L) {
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < PER_T; ++u) {
                const int e = tid * PER_T + u;
                cnt += (e < nmine && lvl[e] == L);
            }
            int inc = cnt;
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) scan_ws[w] = inc;
            __syncthreads();
            int wpre = 0, total = 0;
            for (int ww = 0; ww < GATHER_THREADS / 32; ++ww) {
                const int v = scan_ws[ww];
                if (ww < w) wpre += v;
                total += v;
            }
            int pos = nlist + wpre + inc - cnt;
#pragma unroll
            for (int u = 0; u < PER_T; ++u) {
                const int e = tid * PER_T + u;
                if (e < nmine && lvl[e] == L) list[pos++] = (uint16_t)e;
            }
            if (tid == 0) scan_ws[8 + L] = nlist + total;  // end of level L in the list
            nlist += total;
            __syncthreads();
        }
    }

    float tot[SUBS_PER_THREAD][6];
#pragma unroll
    for (int r = 0; r < SUBS_PER_THREAD; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) tot[r][c] = 0.f;

    auto issue = [&](int e, int stage) {
        const int tile = tbeg + list[e] * tstep;
        const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
        mbar_arrive_expect_tx(&bars[stage], tile_texels * CELL_FLOATS * sizeof(float));
        tma_load_3d(raw0 + stage * RAW_FLOATS, &tmap, &bars[stage], tx * row_floats, ty * tt, rc.env);
    };
    if (g.use_tma && tid == 0) {
        if (0 < nlist) issue(0, 0);
        if (1 < nlist) issue(1, 1);
    }

    // ---- per-level state of my 4 slots: lattice node, texel subset, normals -----------------------------------------
    // slot q = r*256 + tid -> cell q / G, sub = q % G (G slots per cell, a multiple of S^2).  At a level with an
    // Sk x Sk lattice the G slots of a cell form Sk^2 groups of gk = G / Sk^2 slots: the group evaluates one lattice
    // node, its gk slots split the tile's texels (t = u, u+gk, ...).  G divides 256, so the 4 slots of a thread share
    // node and subset.
    float nx[SUBS_PER_THREAD], ny[SUBS_PER_THREAD], nz[SUBS_PER_THREAD], nv[SUBS_PER_THREAD];
    float Fi[SUBS_PER_THREAD], mult[SUBS_PER_THREAD], wq[SUBS_PER_THREAD];
    int cur_level = 0, level_end = 0, gk = 1, u0 = 0;
    const bool hier = (GATHER_THREADS % G) == 0;

    for (int it = 0; it < nlist; ++it) {
        if (it >= level_end) {
            do { ++cur_level; level_end = scan_ws[8 + cur_level]; } while (it >= level_end);  // next non-empty level
            const int Sk = g.lev_S[cur_level - 1];
            gk = hier ? G / (Sk * Sk) : 1;
#pragma unroll
            for (int r = 0; r < SUBS_PER_THREAD; ++r) {
                const int q = r * GATHER_THREADS + tid;
                const int pl = q / G, sub = q - pl * G;
                const int li = pl / g.tile_w, lj = pl - li * g.tile_w;
                const int i = pi0 + li, j = pj0 + lj;
                const int node = sub / gk;
                if (r == 0) u0 = sub - node * gk;
                const int a = node / Sk, b = node - a * Sk;
                const bool active = pl < npix && i < g.res && j < g.res;
                const bool avg = SPEC && g.fine_S > Sk && active;
                float3 va = make_float3(0.f, 0.f, 0.f);
                if (avg)
                    va = view_term_avg(rc, g.cell, i, j, a, b, g.fine_S / Sk, [&](int n) { return g.fine_x[n]; },
                                       [&](int n) { return g.fine_w[n]; });
                if (DIFF) { va.y = 0.f; va.z = 0.f; }  // the diffuse lobe shares the node and has no such factor
                const float th = ((float)i + 0.5f + 0.5f * (g.gl_x[cur_level - 1][a] + va.y)) * g.cell;
                const float ph = ((float)j + 0.5f + 0.5f * (g.gl_x[cur_level - 1][b] + va.z)) * g.cell;
                float st, ct, sp, cp;
                sincosf(th, &st, &ct);
                sincosf(ph, &sp, &cp);
                const float lx = st * cp, lz = st * sp;
                nx[r] = lx * rc.left[0] + ct * rc.upp[0] + lz * rc.vhat[0];
                ny[r] = lx * rc.left[1] + ct * rc.upp[1] + lz * rc.vhat[1];
                nz[r] = lx * rc.left[2] + ct * rc.upp[2] + lz * rc.vhat[2];
                nv[r] = lz;  // n . v exactly, the frame is orthonormal
                const float mm = fminf(fmaxf(1.f - lz, 0.f), 1.f);
                Fi[r] = (mm * mm) * (mm * mm) * mm;
                wq[r] = active ? g.gl_w[cur_level - 1][a] * g.gl_w[cur_level - 1][b] : 0.f;
                // F D G1(n.v) G1(n.d) / (4 n.v) = F x / (q^2 (x + sq)) / (pi alpha^2 (n.v + sqrt((n.v)^2 (1-a^2) + a^2)))
                const float g1 = lz + sqrtf(lz * lz * rc.one_m_a2 + rc.alpha2);
                mult[r] = lz > 0.f ? wq[r] / (3.14159265358979f * rc.alpha2 * g1) : 0.f;
                if (avg) mult[r] = wq[r] * va.x;
            }
        }
        const int stage = it & 1;
        float* raw = raw0 + stage * RAW_FLOATS;
        const int tile = tbeg + list[it] * tstep;
        const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
        if (g.use_tma) {
            mbar_wait(&bars[stage], (it >> 1) & 1);
        } else {
            // plain-load staging for maps whose row pitch is not a multiple of 16 bytes
            const int map_row = g.Wm * CELL_FLOATS;
            const float* src = g.src + (size_t)rc.env * g.Hm * map_row;
            for (int e = tid; e < tile_texels * CELL_FLOATS; e += GATHER_THREADS) {
                const int lr = e / row_floats, lc = e - lr * row_floats;
                const int r = ty * tt + lr, c = tx * row_floats + lc;
                raw[e] = (r < g.Hm && c < map_row) ? src[(size_t)r * map_row + c] : 0.f;
            }
            __syncthreads();
        }

        // ---- transform: raw tile -> pixel-independent records {h, |v+d|, Rr, E dOmega F_c, E dOmega} -----------------
        for (int t = tid; t < tile_texels; t += GATHER_THREADS) {
            const int lr = t / tt, lc = t - lr * tt;
            float dx, dy, dz, er, eg, eb;
            if (COARSE_SRC) {
                const float* cellp = raw + lr * row_floats + lc * COARSE_FLOATS;
                dx = cellp[0]; dy = cellp[1]; dz = cellp[2];
                er = cellp[3]; eg = cellp[4]; eb = cellp[5];
                if (g.far_mode == 2) {
                    // the raw tile this 2x2 cell lies in decides, with the function and inputs of the raw-map launch
                    const int rr = (ty * tt + lr) * g.far_factor, cc = (tx * tt + lc) * g.far_factor;
                    const int rtile = (rr / g.raw_tt) * g.raw_ttiles_x + cc / g.raw_tt;
                    const bool inside = rr < g.raw_Hm && cc < g.raw_Wm;
                    const float dr = inside ? tile_distance_geom(g.cull, g.raw_tt, g.raw_ttiles_x, g.raw_Hm, g.raw_Wm, g.raw_dth,
                                                                 g.raw_dph, g.rc[k].vhat, rtile, ax, ay, az, beta)
                                            : -1.f;
                    if (!far_in_range(g, rc, dr)) { er = 0.f; eg = 0.f; eb = 0.f; }
                }
            } else {
                const int r = min(ty * tt + lr, g.He - 1), c = min(tx * tt + lc, g.We - 1);
                const float st = g.sin_t[r], ct = g.cos_t[r], sp = g.sin_p[c], cp = g.cos_p[c];
                dx = st * sp; dy = ct; dz = -st * cp;
                const float dom = g.domega_k * st;
                er = raw[lr * row_floats + lc * 3 + 0] * dom;
                eg = raw[lr * row_floats + lc * 3 + 1] * dom;
                eb = raw[lr * row_floats + lc * 3 + 2] * dom;
            }
            // |v + d|^2 from its components: 2 + 2 v.d cancels at grazing reflection (d ~ -v) and cost 3 % at the limb cells
            const float sx = rc.vhat[0] + dx, sy = rc.vhat[1] + dy, sz = rc.vhat[2] + dz;
            const float len2 = fmaxf(sx * sx + sy * sy + sz * sz, 1e-12f);
            const float inv_len = rsqrtf(len2);
            const float len = len2 * inv_len;
            const float vh = 0.5f * len;
            rec[t * 3 + 0] = make_float4(sx * inv_len, sy * inv_len, sz * inv_len, len);
            float fr = 0.f, fg = 0.f, fb = 0.f;
            if (SPEC) {
                const float Fd = fresnel_dielectric(vh, rc.eta);
                const float mm = fminf(fmaxf(1.f - vh, 0.f), 1.f);
                const float sw = (mm * mm) * (mm * mm) * mm;
                fr = (1.f - rc.m) * Fd + rc.m * (rc.base[0] + (1.f - rc.base[0]) * sw);
                fg = (1.f - rc.m) * Fd + rc.m * (rc.base[1] + (1.f - rc.base[1]) * sw);
                fb = (1.f - rc.m) * Fd + rc.m * (rc.base[2] + (1.f - rc.base[2]) * sw);
            }
            rec[t * 3 + 1] = make_float4(2.f * rc.rough * vh * vh, er * fr, eg * fg, eb * fb);
            if (DIFF) rec[t * 3 + 2] = make_float4(er, eg, eb, 0.f);
        }
        __syncthreads();  // records ready, raw[stage] free

        if (g.use_tma && tid == 0 && it + 2 < nlist) {
            fence_proxy_async();
            issue(it + 2, stage);
        }

        // ---- gather: my 4 slots x my share of the tile's records ------------------------------------------------
        float acc[SUBS_PER_THREAD][6];
#pragma unroll
        for (int r = 0; r < SUBS_PER_THREAD; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) acc[r][c] = 0.f;

#pragma unroll 2
        for (int t = u0; t < tile_texels; t += gk) {
            const float4 h = rec[t * 3 + 0];
            const float4 s = rec[t * 3 + 1];
            float4 d4;
            if (DIFF) d4 = rec[t * 3 + 2];
#pragma unroll
            for (int r = 0; r < SUBS_PER_THREAD; ++r) {
                const float ex = nx[r] - h.x, ey = ny[r] - h.y, ez = nz[r] - h.z;
                const float u2 = ex * ex + ey * ey + ez * ez;         // 2 (1 - n.h), no cancellation
                const float nh = 1.f - 0.5f * u2;
                const float x = h.w * nh - nv[r];                     // n.d = |v+d| n.h - n.v
                const float xc = fmaxf(x, 0.f);                       // below the horizon: weight 0, denominator > 0
                if (SPEC) {
                    const float sin2 = u2 * (1.f - 0.25f * u2);       // 1 - (n.h)^2
                    const float q = 1.f + sin2 * rc.inv_a2m1;         // cos^2 + sin^2 / alpha^2
                    const float sq = fast_sqrt(xc * xc * rc.one_m_a2 + rc.alpha2);
                    float ws = xc * fast_rcp(q * q * (xc + sq));      // D G1(n.d) up to per-slot constants
                    if (NEAR_EXCL) ws = u2 >= rc.tk2[0] ? ws : 0.f;   // the near field belongs to render_near_kernel
                    acc[r][0] += ws * s.y;
                    acc[r][1] += ws * s.z;
                    acc[r][2] += ws * s.w;
                }
                if (DIFF) {
                    const float mm = 1.f - xc;
                    const float m2 = mm * mm;
                    const float Fo = m2 * m2 * mm;
                    const float Rr = s.x;
                    const float inner = (-0.5f + 0.25f * Fi[r]) + Rr * ((1.f - Fi[r]) + Rr * Fi[r]);
                    const float f = ((1.f - 0.5f * Fi[r]) + Rr * Fi[r]) + Fo * inner;
                    const float wd = xc * f;
                    acc[r][3] += wd * d4.x;
                    acc[r][4] += wd * d4.y;
                    acc[r][5] += wd * d4.z;
                }
            }
        }
        // two-level summation (per tile, then total) keeps the fp32 error near 1e-6; the per-slot constants and the
        // Gauss-Legendre weight of the current level are applied here, once per tile
#pragma unroll
        for (int r = 0; r < SUBS_PER_THREAD; ++r) {
            if (SPEC) {
#pragma unroll
                for (int c = 0; c < 3; ++c) tot[r][c] += mult[r] * acc[r][c];
            }
            if (DIFF) {
#pragma unroll
                for (int c = 3; c < 6; ++c) tot[r][c] += wq[r] * acc[r][c];
            }
        }
        __syncthreads();  // records free
    }

    // ---- epilogue: per-pixel reduction over the S^2 slots in fixed order -----------------------------------------------
    float* resbuf = reinterpret_cast<float*>(rec);  // [SLOTS][3]
#pragma unroll
    for (int r = 0; r < SUBS_PER_THREAD; ++r) {
        const int q = r * GATHER_THREADS + tid;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = SPEC ? tot[r][c] : 0.f;
            if (DIFF) v += rc.cdiff[c] * tot[r][3 + c];
            resbuf[q * 3 + c] = v;
        }
    }
    __syncthreads();
    for (int o = tid; o < npix * 3; o += GATHER_THREADS) {
        const int pl = o / 3, c = o - pl * 3;
        const int li = pl / g.tile_w, lj = pl - li * g.tile_w;
        const int i = pi0 + li, j = pj0 + lj;
        if (i >= g.res || j >= g.res) continue;
        float v = 0.f;
        for (int s2 = 0; s2 < G; ++s2) v += resbuf[(pl * G + s2) * 3 + c];
        const size_t pix = (size_t)i * g.res + j;
        g.slab[(((size_t)blockIdx.z * g.N + k) * g.res * g.res + pix) * 3 + c] = v;
    }
}


// ---- near field of the specular lobe, one CTA per refmap cell ------------------------------------------------------
// The tile kernel's block/tile granularity inflates the region that runs on the fine lattices by an order of magnitude
// for sharp lobes.  Here the split is per (cell, texel): a texel is "near" when the squared chord between the cell's
// centre normal and its half vector is below tk2[0]; the tile kernel (NEAR_EXCL) skips exactly those pairs, this kernel
// evaluates them on the lattice their distance asks for (2x2 ... SxS).  Eight warps scan the rows of the window of
// texels around the mirror direction (coalesced loads straight from the envmap, L2 hits), sort near texels by level
// into per-warp staging lists in shared memory and, whenever a list holds 32 entries, evaluate them lane-per-texel
// against that level's lattice nodes (broadcast LDS); partial sums meet in a warp-shuffle tree and a fixed-order
// reduction over the warps (deterministic).
static constexpr int NEAR_THREADS = 256, NEAR_WARPS = NEAR_THREADS / 32, NEAR_LIST = 64, NEAR_LEVELS = 4;
static constexpr int NEAR_NODES = 4 + 16 + 64 + 256;

struct NearArgs {
    const float* env;
    const RenderConst* rc;
    const float *sin_t, *cos_t, *sin_p, *cos_p;
    float* slab;  // [N][res*res][3]
    int He, We, N, res, nlev;  // nlev lattices: 2, 4, ... 2^nlev
    int view_avg;
    float domega_k, cell;
    float gl_x[NEAR_LEVELS][16], gl_w[NEAR_LEVELS][16];
};

__global__ void __launch_bounds__(NEAR_THREADS) render_near_kernel(const NearArgs g) {
    extern __shared__ __align__(16) unsigned char near_smem[];
    float4* node_n = reinterpret_cast<float4*>(near_smem);                       // [NEAR_NODES] normal, n.v
    float* node_m = reinterpret_cast<float*>(near_smem + NEAR_NODES * 16);       // [NEAR_NODES] weight * constants
    float4* lists = reinterpret_cast<float4*>(near_smem + NEAR_NODES * 20);      // [warp][level][NEAR_LIST][2]
    float* wsum = reinterpret_cast<float*>(near_smem + NEAR_NODES * 20 + NEAR_WARPS * NEAR_LEVELS * NEAR_LIST * 32);

    const int k = blockIdx.y;
    const RenderConst rc = g.rc[k];
    if (!(rc.route & ROUTE_SPEC_RAW)) return;
    const int pix = blockIdx.x;
    const int pi = pix / g.res, pj = pix - pi * g.res;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // lattice nodes of every level: normal, n.v, Gauss-Legendre weight x per-node constants
    int node_off[NEAR_LEVELS + 1];
    node_off[0] = 0;
#pragma unroll
    for (int l = 0; l < NEAR_LEVELS; ++l) node_off[l + 1] = node_off[l] + (l < g.nlev ? (4 << (2 * l)) : 0);
    for (int e = tid; e < node_off[g.nlev]; e += NEAR_THREADS) {
        int l = 0;
        while (e >= node_off[l + 1]) ++l;
        const int Sk = 2 << l, node = e - node_off[l];
        const int a = node / Sk, b = node - a * Sk;
        const bool avg = g.view_avg && l < g.nlev - 1;
        float3 va = make_float3(0.f, 0.f, 0.f);
        if (avg)
            va = view_term_avg(rc, g.cell, pi, pj, a, b, (2 << (g.nlev - 1)) / Sk,
                               [&](int n) { return g.gl_x[g.nlev - 1][n]; }, [&](int n) { return g.gl_w[g.nlev - 1][n]; });
        const float th = ((float)pi + 0.5f + 0.5f * (g.gl_x[l][a] + va.y)) * g.cell;
        const float ph = ((float)pj + 0.5f + 0.5f * (g.gl_x[l][b] + va.z)) * g.cell;
        float st, ct, sp, cp;
        sincosf(th, &st, &ct);
        sincosf(ph, &sp, &cp);
        const float lx = st * cp, lz = st * sp;
        node_n[e] = make_float4(lx * rc.left[0] + ct * rc.upp[0] + lz * rc.vhat[0],
                                lx * rc.left[1] + ct * rc.upp[1] + lz * rc.vhat[1],
                                lx * rc.left[2] + ct * rc.upp[2] + lz * rc.vhat[2], lz);
        const float g1 = lz + sqrtf(lz * lz * rc.one_m_a2 + rc.alpha2);
        node_m[e] = lz > 0.f ? g.gl_w[l][a] * g.gl_w[l][b] / (3.14159265358979f * rc.alpha2 * g1) : 0.f;
        if (avg) node_m[e] = g.gl_w[l][a] * g.gl_w[l][b] * va.x;
    }
    // centre normal: the node of the tile kernel's 1x1 lattice, view-term shift included (view_term_avg with m = S, one
    // fine node per thread and a fixed-order reduction; it may differ from the tile kernel's value in the last bits,
    // which moves a texel across the near / far boundary only if its distance equals the threshold to 1e-7)
    float cnx, cny, cnz;
    {
        __shared__ float cred[NEAR_WARPS][6];
        float3 va = make_float3(0.f, 0.f, 0.f);
        if (g.view_avg) {
            const int m = 2 << (g.nlev - 1);
            float p[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (tid < m * m) {
                const int ia = tid / m, ib = tid - ia * m;
                const float xa = g.gl_x[g.nlev - 1][ia], xb = g.gl_x[g.nlev - 1][ib];
                const float w = g.gl_w[g.nlev - 1][ia] * g.gl_w[g.nlev - 1][ib];
                const float wv = w * view_term(rc, sinf(((float)pi + 0.5f + 0.5f * xa) * g.cell) *
                                                       sinf(((float)pj + 0.5f + 0.5f * xb) * g.cell));
                p[0] = wv; p[1] = w; p[2] = wv * xa; p[3] = wv * xb; p[4] = w * xa; p[5] = w * xb;
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                for (int d = 16; d; d >>= 1) p[c] += __shfl_xor_sync(0xffffffffu, p[c], d);
                if (lane == 0) cred[warp][c] = p[c];
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                p[c] = 0.f;
                for (int w = 0; w < NEAR_WARPS; ++w) p[c] += cred[w][c];
            }
            if (p[0] > 0.f) va = make_float3(p[0] / p[1], p[2] / p[0] - p[4] / p[1], p[3] / p[0] - p[5] / p[1]);
        }
        const float th = ((float)pi + 0.5f + 0.5f * (0.f + va.y)) * g.cell, ph = ((float)pj + 0.5f + 0.5f * (0.f + va.z)) * g.cell;
        float st, ct, sp, cp;
        sincosf(th, &st, &ct);
        sincosf(ph, &sp, &cp);
        const float lx = st * cp, lz = st * sp;
        cnx = lx * rc.left[0] + ct * rc.upp[0] + lz * rc.vhat[0];
        cny = lx * rc.left[1] + ct * rc.upp[1] + lz * rc.vhat[1];
        cnz = lx * rc.left[2] + ct * rc.upp[2] + lz * rc.vhat[2];
    }
    __syncthreads();

    // window of texels that can be near: within 2 * angle(T0) of the mirror direction (reflection doubles angles)
    const float cv = cnx * rc.vhat[0] + cny * rc.vhat[1] + cnz * rc.vhat[2];
    const float rx = 2.f * cv * cnx - rc.vhat[0], ry = 2.f * cv * cny - rc.vhat[1], rz = 2.f * cv * cnz - rc.vhat[2];
    const float dth = 3.14159265f / g.He, dph = 6.2831853f / g.We;
    const float chord0 = sqrtf(rc.tk2[0]);
    const float Rd = 4.f * asinf(fminf(1.f, 0.5f * chord0)) + 2.f * (dth + dph) + 0.01f;
    const float th_r = acosf(fminf(fmaxf(ry, -1.f), 1.f));
    float ph_r = atan2f(rx, -rz);
    if (ph_r < 0.f) ph_r += 6.2831853f;
    const bool whole = Rd >= 3.1f;
    const int r_lo = whole ? 0 : max(0, (int)floorf((th_r - Rd) / dth));
    const int r_hi = whole ? g.He - 1 : min(g.He - 1, (int)ceilf((th_r + Rd) / dth));
    const float cosR = cosf(fminf(Rd, 3.14159265f)), sr = sinf(th_r), cr = cosf(th_r);

    // staging lists of this warp, half vectors and radiances in separate arrays (16-byte stride: conflict-free LDS/STS)
    float4* hlist = lists + (size_t)warp * NEAR_LEVELS * NEAR_LIST;
    float4* elist = lists + (size_t)(NEAR_WARPS + warp) * NEAR_LEVELS * NEAR_LIST;
    int cnt[NEAR_LEVELS] = {0, 0, 0, 0};
    float tot0 = 0.f, tot1 = 0.f, tot2 = 0.f;
    const float* src = g.env + (size_t)rc.env * g.He * g.We * 3;

    auto process = [&](int l, int n) {  // lanes < n evaluate their staged texel against every node of lattice l
        const float4 h = hlist[l * NEAR_LIST + (lane < n ? lane : 0)];
        const float4 es = elist[l * NEAR_LIST + (lane < n ? lane : 0)];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const int n0 = node_off[l], n1 = node_off[l + 1];
        for (int j = n0; j < n1; ++j) {
            const float4 nn = node_n[j];
            const float mlt = node_m[j];
            const float ex = nn.x - h.x, ey = nn.y - h.y, ez = nn.z - h.z;
            const float u2 = ex * ex + ey * ey + ez * ez;
            const float nh = 1.f - 0.5f * u2;
            const float x = h.w * nh - nn.w;
            const float xc = fmaxf(x, 0.f);
            const float sin2 = u2 * (1.f - 0.25f * u2);
            const float q = 1.f + sin2 * rc.inv_a2m1;
            const float sq = fast_sqrt(xc * xc * rc.one_m_a2 + rc.alpha2);
            const float ws = mlt * (xc * fast_rcp(q * q * (xc + sq)));
            a0 += ws * es.x;
            a1 += ws * es.y;
            a2 += ws * es.z;
        }
        if (lane < n) { tot0 += a0; tot1 += a1; tot2 += a2; }
    };

    for (int r = r_lo + warp; r <= r_hi; r += NEAR_WARPS) {
        const float st = g.sin_t[r], ct = g.cos_t[r];
        int c_start = 0, ncols = g.We;
        if (!whole) {
            const float num = cosR - cr * ct, den = sr * st;
            if (den > 1e-6f && num > -den) {
                if (num >= den) continue;  // the row does not reach the window
                const float dphi = acosf(num / den);
                ncols = min(g.We, 2 * (int)ceilf(dphi / dph) + 3);
                c_start = (int)floorf(ph_r / dph) - ncols / 2;
                c_start = ((c_start % g.We) + g.We) % g.We;
            }
        }
        const float dom = g.domega_k * st;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            const bool in_row = c0 + lane < ncols;
            int c = c_start + c0 + lane;
            if (c >= g.We) c -= g.We;
            int lvl = -1;
            float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f), e4 = h4;
            if (in_row) {
                const float sp = g.sin_p[c], cp = g.cos_p[c];
                const float dx = st * sp, dy = ct, dz = -st * cp;
                const float sx = rc.vhat[0] + dx, sy = rc.vhat[1] + dy, sz = rc.vhat[2] + dz;
                const float len2 = fmaxf(sx * sx + sy * sy + sz * sz, 1e-12f);
                const float inv_len = rsqrtf(len2);
                const float len = len2 * inv_len;
                h4 = make_float4(sx * inv_len, sy * inv_len, sz * inv_len, len);
                const float ex = cnx - h4.x, ey = cny - h4.y, ez = cnz - h4.z;
                const float u2 = ex * ex + ey * ey + ez * ez;
                if (u2 < rc.tk2[0]) {
                    lvl = g.nlev - 1;  // lattice 2^(l+1): the coarsest one whose threshold the distance passes
                    for (int l = 0; l < g.nlev - 1; ++l)
                        if (u2 >= rc.tk2[l + 1]) { lvl = l; break; }
                    const float vh = 0.5f * len;
                    const float Fd = fresnel_dielectric(vh, rc.eta);
                    const float mm = fminf(fmaxf(1.f - vh, 0.f), 1.f);
                    const float sw = (mm * mm) * (mm * mm) * mm;
                    const float* e = src + ((size_t)r * g.We + c) * 3;
                    e4.x = e[0] * dom * ((1.f - rc.m) * Fd + rc.m * (rc.base[0] + (1.f - rc.base[0]) * sw));
                    e4.y = e[1] * dom * ((1.f - rc.m) * Fd + rc.m * (rc.base[1] + (1.f - rc.base[1]) * sw));
                    e4.z = e[2] * dom * ((1.f - rc.m) * Fd + rc.m * (rc.base[2] + (1.f - rc.base[2]) * sw));
                }
            }
#pragma unroll
            for (int l = 0; l < NEAR_LEVELS; ++l) {
                if (l >= g.nlev) break;
                const unsigned m = __ballot_sync(0xffffffffu, lvl == l);
                if (!m) continue;
                if (lvl == l) {
                    const int pos = cnt[l] + __popc(m & ((1u << lane) - 1u));
                    hlist[l * NEAR_LIST + pos] = h4;
                    elist[l * NEAR_LIST + pos] = e4;
                }
                cnt[l] += __popc(m);
                __syncwarp();
                if (cnt[l] >= 32) {
                    process(l, 32);
                    // move the overflow (entries 32 .. cnt-1) to the front
                    const int rest = cnt[l] - 32;
                    float4 t0, t1;
                    if (lane < rest) {
                        t0 = hlist[l * NEAR_LIST + 32 + lane];
                        t1 = elist[l * NEAR_LIST + 32 + lane];
                    }
                    __syncwarp();
                    if (lane < rest) {
                        hlist[l * NEAR_LIST + lane] = t0;
                        elist[l * NEAR_LIST + lane] = t1;
                    }
                    cnt[l] = rest;
                    __syncwarp();
                }
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NEAR_LEVELS; ++l)
        if (l < g.nlev && cnt[l] > 0) process(l, cnt[l]);

    for (int d = 16; d; d >>= 1) {
        tot0 += __shfl_xor_sync(0xffffffffu, tot0, d);
        tot1 += __shfl_xor_sync(0xffffffffu, tot1, d);
        tot2 += __shfl_xor_sync(0xffffffffu, tot2, d);
    }
    if (lane == 0) { wsum[warp * 3 + 0] = tot0; wsum[warp * 3 + 1] = tot1; wsum[warp * 3 + 2] = tot2; }
    __syncthreads();
    if (tid < 3) {
        float v = 0.f;
        for (int w = 0; w < NEAR_WARPS; ++w) v += wsum[w * 3 + tid];
        g.slab[((size_t)k * g.res * g.res + pix) * 3 + tid] = v;
    }
}

struct SlabDesc {
    const float* base;
    int splits, route_mask;
    float sign;  // +1, or -1 for the centre-sample half of the diffuse lattice correction
};

static constexpr int MAX_SLABS = 10;
struct SlabSet {
    SlabDesc s[MAX_SLABS];
};

// out = sum over the launches that served the render and over their texel splits, in fixed order (deterministic)
__global__ void render_combine_kernel(const SlabSet set, const RenderConst* __restrict__ rc, float* __restrict__ out,
                                      int N, int res, int channel_first) {
    const size_t total = (size_t)N * res * res * 3;
    size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int c = (int)(o % 3);
    const size_t pix = (o / 3) % ((size_t)res * res);
    const size_t k = o / 3 / ((size_t)res * res);
    const int route = rc[k].route;
    float v = 0.f;
#pragma unroll
    for (int l = 0; l < MAX_SLABS; ++l)
        if (set.s[l].base && (route & set.s[l].route_mask))
            for (int sp = 0; sp < set.s[l].splits; ++sp) v += set.s[l].sign * set.s[l].base[(size_t)sp * total + o];
    const size_t idx = channel_first ? (k * 3 + c) * res * res + pix : o;
    out[idx] = v;
}

struct RenderPlan {
    int S, G, tt, tile_w, tile_h, tiles_x, tiles_y, ttiles_x, ttiles_y, splits;
};

// tile edge: tiles of about 3 degrees give the level schedule something to decide (measured at 1000 rows: 16 texels
// 43.9 refmaps/s, 32 texels 42.3, 8 texels 37.9, where the per-tile overhead takes over)
static int tile_edge(int Hm) { return Hm >= 1600 ? 32 : Hm >= 400 ? 16 : 8; }

// slots per cell: S^2, raised to 16 for the 2x2 and 4x4 lattices so that their blocks are 8x8 cells (a tighter cone of
// normals sends more tiles to the coarse levels) and to 4 for the 1x1 lattice (16x16-cell blocks cull a third of the
// tiles that 32x32-cell blocks would visit)
static int default_slots(int S) { return S == 1 ? 4 : (S == 2 || S == 4) ? 16 : S * S; }

// slots per cell of the raw-map launch and its coarse-map twins (they must share the block of cells)
static int raw_slots(int S) {
    const char* e = getenv(S == 2 ? "DRM_RENDER_SLOTS2" : S == 4 ? "DRM_RENDER_SLOTS4" : "DRM_RENDER_SLOTS_NONE");
    if (e) { const int v = atoi(e); if (v == 16 || v == 64 || v == 256) return v; }
    // 4x4 footprints: 4x4-cell blocks (tighter cones) measured 2 % faster than 8x8 once the far field left the raw map;
    // 2x2 footprints: 8x8-cell blocks stay ahead
    return S == 4 ? 64 : default_slots(S);
}

static RenderPlan make_plan(int N, int Hm, int Wm, int res, int S, int G, int tt) {
    RenderPlan p;
    p.S = S;
    p.G = G;
    int px = SLOTS / G;
    int e = (int)floor(sqrt((double)px));
    if (e < 1) e = 1;
    if (e > res) e = res;
    p.tile_w = p.tile_h = e;
    p.tiles_x = (res + e - 1) / e;
    p.tiles_y = (res + e - 1) / e;
    p.tt = tt;
    p.ttiles_x = (Wm + tt - 1) / tt;
    p.ttiles_y = (Hm + tt - 1) / tt;
    const long ctas = (long)p.tiles_x * p.tiles_y * N;
    // twelve waves of two resident CTAs per SM: a launch may serve only the fraction of the N renders routed to it (the
    // routing is decided on the device), so the grid is over-split to keep the GPU full and balanced in that case
    const long want = 148L * 2 * 12;
    long s = (want + ctas - 1) / ctas;
    const long ntiles = (long)p.ttiles_x * p.ttiles_y;
    const long smax = ntiles / 4 > 0 ? ntiles / 4 : 1;
    if (s > smax) s = smax;
    const long smin = (ntiles + MAX_LIST - 1) / MAX_LIST;  // a CTA schedules at most MAX_LIST tiles
    if (s < smin) s = smin;
    if (s < 1) s = 1;
    p.splits = (int)s;
    return p;
}

struct RenderLayout {
    int Hc, Wc, Hc2, Wc2, Hc16, Wc16;
    bool coarse_enabled, coarse_diffuse_ok;
    float coarse_h, coarse2_h;
    bool far_pair;  // the raw-map launch is split into a far launch (16x16-cell blocks) and a near launch
    bool pow2;      // S in {2,4,8,16}: the footprint hierarchy (and the per-cell near-field kernel) applies
    RenderPlan raw, far, diff, coarse, coarse2;  // launches: spec/both on the raw map (+ its far half), diffuse / both on the coarse maps
    RenderPlan mid, farc;  // near-mode far field: raw map within dfar / 2x2 coarse map beyond, both with 8x8-cell blocks
    RenderPlan farc_raw;   // the same split for the block/tile schedule: coarse-map twin of `raw`
    RenderPlan farc4, farc4_raw;  // ... and beyond dfar4 the 4x4 coarse map
    bool far_coarse4;
    float dfar4;
    RenderPlan diff1, corr_hi, corr_lo;  // diffuse lobe: 1x1 lattice on the 4x4 map + (lattice - centre) on the 16x16 map
    bool diff_corr;
    bool far_coarse_any;
    bool far_coarse;       // the far field of sharp lobes beyond dfar is gathered from the 2x2 coarse map
    float dfar;
    RenderConst* rc;
    int* env_used;
    float *sin_t, *cos_t, *sin_p, *cos_p, *coarse_map, *coarse2_map, *slab_raw, *slab_far, *slab_near, *slab_diff, *slab_coarse,
        *slab_coarse2, *slab_farc, *slab_farc4, *coarse16_map, *slab_corr_hi, *slab_corr_lo;
};

static size_t render_layout(RenderLayout& L, void* ws, int N, int B, int He, int We, int res, int S) {
    L.Hc = (He + COARSE - 1) / COARSE;
    L.Wc = (We + COARSE - 1) / COARSE;
    L.Hc2 = (He + COARSE2 - 1) / COARSE2;
    L.Wc2 = (We + COARSE2 - 1) / COARSE2;
    L.Hc16 = (L.Hc + 3) / 4;
    L.Wc16 = (L.Wc + 3) / 4;
    L.coarse_h = (float)(COARSE * M_PI / He);
    L.coarse2_h = (float)(COARSE2 * M_PI / He);
    const char* cv = getenv("DRM_RENDER_COARSE");  // "0" disables the coarse-map routes (debugging / validation)
    L.coarse_enabled = !(cv && cv[0] == '0') && He >= 8 * COARSE && We >= 8 * COARSE;
    L.coarse_diffuse_ok = L.coarse_enabled && L.coarse_h <= 0.0135f;
    L.pow2 = (S == 2 || S == 4 || S == 8 || S == 16);
    L.far_pair = (S == 8 || S == 16) && res >= FAR_EDGE;
    L.raw = make_plan(N, He, We, res, S, raw_slots(S), tile_edge(He));
    L.far = make_plan(N, He, We, res, 1, SLOTS / (FAR_EDGE * FAR_EDGE), tile_edge(He));
    // the coarse maps keep 32x32-cell tiles: their launches run few lattice levels and gain nothing from finer tiles
    L.diff = make_plan(N, L.Hc, L.Wc, res, S < 2 ? S : 2, default_slots(S < 2 ? S : 2), TT);
    // S >= 2: the cell average of the diffuse lobe differs from its centre sample by ~5e-5; that difference is smooth in
    // the light direction, so it is taken from a 16x16-texel coarsening while the centre sample keeps the 4x4 map
    const char* dc = getenv("DRM_RENDER_DIFF_CORR");  // "0" disables (debugging / validation)
    L.diff_corr = L.coarse_diffuse_ok && S >= 2 && !(dc && dc[0] == '0');
    L.diff1 = make_plan(N, L.Hc, L.Wc, res, 1, default_slots(1), TT);
    L.corr_hi = make_plan(N, L.Hc16, L.Wc16, res, 2, default_slots(2), TT);
    L.corr_lo = make_plan(N, L.Hc16, L.Wc16, res, 1, default_slots(1), TT);
    L.coarse = make_plan(N, L.Hc, L.Wc, res, S, default_slots(S), TT);
    L.coarse2 = make_plan(N, L.Hc2, L.Wc2, res, S, default_slots(S), TT);
    // Far field of sharp lobes (S >= 8) from the 2x2 coarse map: a tail ~ d^-4 evaluated once per cell of half-vector
    // size h = coarse2_h / 2 is off by (h^2 / 24)(20 / d^2) locally; beyond dfar that is below 1e-4 even for a pixel
    // whose value is all halo.
    L.dfar = 0.5f * L.coarse2_h * sqrtf(20.f / (24.f * 1e-4f));
    if (const char* ds = getenv("DRM_RENDER_DFAR_SCALE")) L.dfar *= (float)atof(ds);
    const char* fc = getenv("DRM_RENDER_FAR_COARSE");  // "0" disables (debugging / validation)
    const char* lv0 = getenv("DRM_RENDER_LEVELS");  // "0": the single-level validation mode evaluates the full sum
    L.far_coarse_any = L.coarse_enabled && L.dfar < 1.2f && !(fc && fc[0] == '0') && !(lv0 && lv0[0] == '0');
    L.far_coarse = L.far_coarse_any && (S == 8 || S == 16);
    L.farc_raw = make_plan(N, L.Hc2, L.Wc2, res, S, raw_slots(S), FARC_TT);
    L.mid = make_plan(N, He, We, res, 1, 16, tile_edge(He));
    L.farc = make_plan(N, L.Hc2, L.Wc2, res, 1, 16, FARC_TT);
    // twice as far the 4x4 map is as accurate as the 2x2 map at dfar
    L.dfar4 = L.dfar * (float)(COARSE / COARSE2);
    const char* f4 = getenv("DRM_RENDER_FAR_COARSE4");  // "0" disables (debugging / validation)
    L.far_coarse4 = L.far_coarse_any && L.dfar4 < 1.4f && !(f4 && f4[0] == '0');
    L.farc4_raw = make_plan(N, L.Hc, L.Wc, res, S, raw_slots(S), FARC_TT);
    L.farc4 = make_plan(N, L.Hc, L.Wc, res, 1, 16, FARC_TT);
    Carver c(ws);
    const size_t slice = (size_t)N * res * res * 3;
    L.rc = c.take<RenderConst>(N);
    L.sin_t = c.take<float>(He);
    L.cos_t = c.take<float>(He);
    L.sin_p = c.take<float>(We);
    L.cos_p = c.take<float>(We);
    L.env_used = c.take<int>(B);
    L.coarse_map = c.take<float>(L.coarse_enabled ? (size_t)B * L.Hc * L.Wc * COARSE_FLOATS : 1);
    L.coarse2_map = c.take<float>(L.coarse_enabled ? (size_t)B * L.Hc2 * L.Wc2 * COARSE_FLOATS : 1);
    L.slab_raw = c.take<float>(slice * L.raw.splits);
    L.slab_far = c.take<float>(L.pow2 ? slice * (L.far.splits > L.mid.splits ? L.far.splits : L.mid.splits) : 1);
    L.slab_near = c.take<float>(L.pow2 ? slice : 1);
    L.slab_diff = c.take<float>(L.coarse_diffuse_ok ? slice * (L.diff.splits > L.diff1.splits ? L.diff.splits : L.diff1.splits) : 1);
    L.coarse16_map = c.take<float>(L.diff_corr ? (size_t)B * L.Hc16 * L.Wc16 * COARSE_FLOATS : 1);
    L.slab_corr_hi = c.take<float>(L.diff_corr ? slice * L.corr_hi.splits : 1);
    L.slab_corr_lo = c.take<float>(L.diff_corr ? slice * L.corr_lo.splits : 1);
    L.slab_coarse = c.take<float>(L.coarse_enabled ? slice * L.coarse.splits : 1);
    L.slab_coarse2 = c.take<float>(L.coarse_enabled ? slice * L.coarse2.splits : 1);
    L.slab_farc = c.take<float>(L.far_coarse_any ? slice * (L.farc.splits > L.farc_raw.splits ? L.farc.splits : L.farc_raw.splits) : 1);
    L.slab_farc4 = c.take<float>(L.far_coarse4 ? slice * (L.farc4.splits > L.farc4_raw.splits ? L.farc4.splits : L.farc4_raw.splits) : 1);
    return c.used();
}

static void gauss_legendre(int S, float* x, float* w) {
    // Newton iteration on P_S; nodes ascending, weights normalised to sum 1
    for (int i = 0; i < S; ++i) {
        double z = cos(M_PI * (i + 0.75) / (S + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < S; ++j) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
            }
            pp = S * (z * p1 - p2) / (z * z - 1.0);
            double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-15) break;
        }
        x[S - 1 - i] = (float)z;
        w[S - 1 - i] = (float)(1.0 / ((1.0 - z * z) * pp * pp));  // = w_i / 2
    }
}

// footprint levels of a launch: power-of-two lattices below S when S is one of 2,4,8,16 (S^2 then divides the 256
// threads); any other S runs as a single level
static void set_levels(GatherArgs& g, int S, bool hierarchy, int fine_S = 0) {
    const char* va = getenv("DRM_RENDER_VIEW_AVG");  // "0": coarse lattice nodes carry their own view term (validation)
    const bool fine_pow2 = (fine_S == 2 || fine_S == 4 || fine_S == 8 || fine_S == 16);
    g.fine_S = (va && va[0] == '0') || !fine_pow2 ? 0 : fine_S;
    if (g.fine_S) gauss_legendre(g.fine_S, g.fine_x, g.fine_w);
    g.S = S;
    g.nlev = 0;
    const bool pow2 = (S == 2 || S == 4 || S == 8 || S == 16);
    if (hierarchy && pow2)
        for (int sk = 1, ti = 0; sk < S; sk *= 2, ++ti) {
            g.lev_S[g.nlev] = sk;
            g.lev_tidx[g.nlev] = ti;
            gauss_legendre(sk, g.gl_x[g.nlev], g.gl_w[g.nlev]);
            ++g.nlev;
        }
    g.lev_S[g.nlev] = S;
    g.lev_tidx[g.nlev] = MAX_LEVELS - 1;
    gauss_legendre(S, g.gl_x[g.nlev], g.gl_w[g.nlev]);
    ++g.nlev;
}

static int make_tensor_map(CUtensorMap* tmap, const float* base, int B, int Hm, int Wm, int floats_per_cell, int tt) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("render: cuTensorMapEncodeTiled entry point unavailable");
        return DRM_ECUDA;
    }
    const cuuint64_t row = (cuuint64_t)Wm * floats_per_cell;
    cuuint64_t dims[3] = {row, (cuuint64_t)Hm, (cuuint64_t)B};
    cuuint64_t strides[2] = {row * 4, row * 4 * (cuuint64_t)Hm};
    cuuint32_t box[3] = {(cuuint32_t)(tt * floats_per_cell), (cuuint32_t)tt, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("render: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return DRM_ECUDA;
    }
    return DRM_OK;
}

template <int TERMS, bool COARSE_SRC, bool NEAR_EXCL = false>
static int launch_gather(const GatherArgs& g, const CUtensorMap& tmap, const RenderPlan& p, int N, cudaStream_t st) {
    const size_t raw_floats = (size_t)TILE_TEXELS * (COARSE_SRC ? COARSE_FLOATS : 3);
    const size_t smem = 2 * raw_floats * sizeof(float) + TILE_TEXELS * REC_FLOATS * sizeof(float) + 16 + 64 + MAX_LIST * 3;
    DRM_CHECK_CUDA(cudaFuncSetAttribute(render_gather_kernel<TERMS, COARSE_SRC, NEAR_EXCL>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(p.tiles_x * p.tiles_y, N, p.splits);
    render_gather_kernel<TERMS, COARSE_SRC, NEAR_EXCL><<<grid, GATHER_THREADS, smem, st>>>(tmap, g);
    count_launches(1);
    return DRM_OK;
}

}  // namespace drm

using namespace drm;

extern "C" size_t drm_render_flat_workspace_bytes(int N, int B, int He, int We, int res, int S) {
    if (N <= 0 || B <= 0 || He <= 0 || We <= 0 || res <= 0 || S < 1 || S > 16) return 0;
    RenderLayout L;
    return render_layout(L, nullptr, N, B, He, We, res, S);
}

extern "C" int drm_render_refmaps_flat(const float* env, int B, int He, int We, const int32_t* env_index, const float* z6,
                                  const float* view3, const uint8_t* flip, int N, int res, int S, float alpha_min,
                                  int channel_first, float* out, void* workspace, size_t workspace_bytes,
                                  void* cuda_stream) {
    DRM_REQUIRE(env && z6 && view3 && out, "render: null pointer");
    DRM_REQUIRE(N > 0 && B > 0 && He > 0 && We > 0 && res > 0, "render: N=%d B=%d He=%d We=%d res=%d must be positive", N, B, He, We, res);
    DRM_REQUIRE(S >= 1 && S <= 16, "render: footprint_S=%d not in 1..16", S);
    DRM_REQUIRE(res <= 4096, "render: res=%d too large", res);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    RenderLayout L;
    const size_t need = render_layout(L, workspace, N, B, He, We, res, S);
    if (!workspace || workspace_bytes < need) {
        set_error("render: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return DRM_EWORKSPACE;
    }
    if (!(alpha_min > 0.f)) alpha_min = fmaxf(1e-3f, (float)(1.25 * M_PI / He));
    const char* lv = getenv("DRM_RENDER_LEVELS");  // "0" disables the footprint hierarchy (debugging / validation)
    const bool hierarchy = !(lv && lv[0] == '0');
    // thresholds of render_setup_kernel are conservative at res 128: 0.3 measured (scripts/levels_probe.py).  Coarser
    // refmaps have cells much wider than the lobe and the grazing-incidence kink of G1(n.d) at the limb is then the
    // binding error (scripts/lowres_probe.py): the thresholds grow with the cell size, which costs little there.
    float level_scale = 0.3f * (float)fmin(10.0, pow(fmax(1.0, (M_PI / res) / (M_PI / 128.0)), 1.5));
    if (const char* ls = getenv("DRM_RENDER_LEVEL_SCALE")) level_scale = (float)atof(ls);

    GatherArgs g{};
    g.rc = L.rc; g.sin_t = L.sin_t; g.cos_t = L.cos_t; g.sin_p = L.sin_p; g.cos_p = L.cos_p;
    g.B = B; g.He = He; g.We = We; g.N = N; g.res = res;
    g.cull = 1;
    g.part = PART_ALL;
    g.far_edge = FAR_EDGE;
    g.domega_k = (float)((2.0 * M_PI / We) * (M_PI / He));
    g.cell = (float)(M_PI / res);

    const char* nvs = getenv("DRM_RENDER_NEAR");
    const char* un = getenv("DRM_RENDER_UNIFY");  // "0": moderately rough renders keep the both-lobes 2x2 launch (validation)
    const bool unify = L.far_coarse4 && L.coarse_diffuse_ok && hierarchy && !(un && un[0] == '0') &&
                       (!L.far_pair || !(nvs && nvs[0] == '0'));
    const int tb = 128;
    render_tables_kernel<<<(max(He, We) + tb - 1) / tb, tb, 0, st>>>(L.sin_t, L.cos_t, L.sin_p, L.cos_p, He, We);
    float near_scale = 2.f * level_scale;
    if (const char* ns = getenv("DRM_RENDER_NEAR_SCALE")) near_scale = (float)atof(ns);
    render_setup_kernel<<<(N + tb - 1) / tb, tb, 0, st>>>(z6, view3, flip, env_index, N, B, alpha_min, g.cell, level_scale, near_scale,
                                                         L.coarse_enabled ? L.coarse_h : 0.f,
                                                         L.coarse_enabled ? L.coarse2_h : 0.f, L.coarse_diffuse_ok ? 1 : 0,
                                                         unify ? 1 : 0, L.rc);
    count_launches(2);
    if (L.coarse_enabled) {
        DRM_CHECK_CUDA(cudaMemsetAsync(L.env_used, 0, sizeof(int) * B, st));
        env_mark_used_kernel<<<(N + tb - 1) / tb, tb, 0, st>>>(L.rc, N, L.env_used);
        const long cells = (long)B * L.Hc * L.Wc;
        env_coarsen_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(env, L.sin_t, L.cos_t, L.sin_p, L.cos_p, L.env_used,
                                                                            B, He, We, L.Hc, L.Wc, L.Hc2, L.Wc2, g.domega_k,
                                                                            L.coarse_map, L.coarse2_map);
        if (L.diff_corr) {
            const long cells16 = (long)B * L.Hc16 * L.Wc16;
            env_coarsen16_kernel<<<(unsigned)((cells16 + 255) / 256), 256, 0, st>>>(L.coarse_map, L.env_used, B, L.Hc, L.Wc,
                                                                                L.Hc16, L.Wc16, L.coarse16_map);
            count_launches(1);
        }
        count_launches(2);
    }

    auto fill_plan = [&](GatherArgs& a, const RenderPlan& p) {
        a.tile_w = p.tile_w; a.tile_h = p.tile_h; a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y;
        a.ttiles_x = p.ttiles_x; a.ttiles_y = p.ttiles_y; a.splits = p.splits;
        a.G = p.G; a.tt = p.tt;
    };
    int rc_code;
    bool used_far = false, near_mode = false, used_farc = false;
    // ---- launches on the raw map: specular lobe only, and both lobes -------------------------------------------------
    {
        GatherArgs a = g;
        a.src = env; a.Hm = He; a.Wm = We; a.slab = L.slab_raw;
        a.dth_cell = (float)(M_PI / He); a.dph_cell = (float)(2.0 * M_PI / We);
        fill_plan(a, L.raw);
        set_levels(a, S, hierarchy, S);
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        a.use_tma = ((We * 12) % 16 == 0) && ((reinterpret_cast<uintptr_t>(env) & 15) == 0);
        if (a.use_tma && (rc_code = make_tensor_map(&tmap, env, B, He, We, 3, L.raw.tt)) != DRM_OK) return rc_code;
        const bool pair = L.far_pair && hierarchy;
        const char* nv = getenv("DRM_RENDER_NEAR");  // "0": block/tile level schedule instead of the per-cell near kernel
        // measured (scripts/levels_probe.py): 2.4x / 1.6x faster than the block schedule for the 16x16 / 8x8 footprints,
        // slower for 4x4 and 2x2 whose windows are too wide to scan once per cell
        near_mode = L.pow2 && S >= 8 && hierarchy && !(nv && nv[0] == '0');
        // far half of a pair / NEAR_EXCL launch: 16x16-cell blocks, 1x1 lattice, 4 slots per cell share each tile
        GatherArgs f = a;
        f.slab = L.slab_far;
        fill_plan(f, L.far);
        set_levels(f, 1, false, S);
        if (near_mode) {
            // specular-only renders: every (cell, texel) pair beyond the 1x1 threshold here, the rest per cell below
            f.part = PART_ALL; f.route_mask = ROUTE_SPEC_RAW;
            used_farc = L.far_coarse;
            if (used_farc) {
                // ... and beyond dfar from the 2x2 coarse map: both launches use 8x8-cell blocks and take the decision
                // per (block, raw tile) from the same function, so every pair is evaluated exactly once
                fill_plan(f, L.mid);
                f.far_mode = 1; f.dfar = L.dfar; f.dfar_near = 1;
                if ((rc_code = launch_gather<1, false, true>(f, tmap, L.mid, N, st)) != DRM_OK) return rc_code;
                GatherArgs c2 = g;
                c2.src = L.coarse2_map; c2.Hm = L.Hc2; c2.Wm = L.Wc2; c2.slab = L.slab_farc;
                c2.dth_cell = (float)(COARSE2 * M_PI / He); c2.dph_cell = (float)(COARSE2 * 2.0 * M_PI / We);
                fill_plan(c2, L.farc);
                set_levels(c2, 1, false, S);
                c2.part = PART_ALL; c2.route_mask = ROUTE_SPEC_RAW;
                c2.far_mode = 2; c2.dfar = L.dfar; c2.dfar_near = 1;
                c2.raw_tt = L.mid.tt; c2.raw_ttiles_x = L.mid.ttiles_x; c2.raw_Hm = He; c2.raw_Wm = We;
                c2.raw_dth = a.dth_cell; c2.raw_dph = a.dph_cell;
                CUtensorMap tmapc;
                memset(&tmapc, 0, sizeof(tmapc));
                c2.use_tma = ((L.Wc2 * COARSE_FLOATS * 4) % 16 == 0);
                if (c2.use_tma && (rc_code = make_tensor_map(&tmapc, L.coarse2_map, B, L.Hc2, L.Wc2, COARSE_FLOATS, FARC_TT)) != DRM_OK) return rc_code;
                c2.far_factor = COARSE2; c2.dfar_hi = L.far_coarse4 ? L.dfar4 : 0.f;
                if ((rc_code = launch_gather<1, true>(c2, tmapc, L.farc, N, st)) != DRM_OK) return rc_code;
                if (L.far_coarse4) {
                    GatherArgs c4 = c2;
                    c4.src = L.coarse_map; c4.Hm = L.Hc; c4.Wm = L.Wc; c4.slab = L.slab_farc4;
                    c4.dth_cell = (float)(COARSE * M_PI / He); c4.dph_cell = (float)(COARSE * 2.0 * M_PI / We);
                    fill_plan(c4, L.farc4);
                    c4.far_factor = COARSE; c4.dfar = L.dfar4; c4.dfar_hi = 0.f;
                    CUtensorMap tmap4;
                    memset(&tmap4, 0, sizeof(tmap4));
                    c4.use_tma = ((L.Wc * COARSE_FLOATS * 4) % 16 == 0);
                    if (c4.use_tma && (rc_code = make_tensor_map(&tmap4, L.coarse_map, B, L.Hc, L.Wc, COARSE_FLOATS, FARC_TT)) != DRM_OK) return rc_code;
                    if ((rc_code = launch_gather<1, true>(c4, tmap4, L.farc4, N, st)) != DRM_OK) return rc_code;
                }
            } else {
                if ((rc_code = launch_gather<1, false, true>(f, tmap, L.far, N, st)) != DRM_OK) return rc_code;
            }
            NearArgs n{};
            n.env = env; n.rc = L.rc; n.sin_t = L.sin_t; n.cos_t = L.cos_t; n.sin_p = L.sin_p; n.cos_p = L.cos_p;
            n.slab = L.slab_near; n.He = He; n.We = We; n.N = N; n.res = res;
            n.domega_k = g.domega_k; n.cell = g.cell;
            n.nlev = 0;
            {
                const char* va = getenv("DRM_RENDER_VIEW_AVG");
                n.view_avg = !(va && va[0] == '0');
            }
            for (int sk = 2; sk <= S; sk *= 2) {
                gauss_legendre(sk, n.gl_x[n.nlev], n.gl_w[n.nlev]);
                ++n.nlev;
            }
            const size_t nsmem = NEAR_NODES * 20 + (size_t)NEAR_WARPS * NEAR_LEVELS * NEAR_LIST * 32 + NEAR_WARPS * 3 * 4;
            DRM_CHECK_CUDA(cudaFuncSetAttribute(render_near_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nsmem));
            render_near_kernel<<<dim3(res * res, N), NEAR_THREADS, nsmem, st>>>(n);
            count_launches(1);
        } else {
            a.part = pair ? PART_NEAR : PART_ALL;
            a.route_mask = ROUTE_SPEC_RAW;
            // same split for the block/tile schedule (any S): tiles beyond dfar come from the 2x2 coarse map, evaluated
            // with the same blocks and lattices (a coarse tile is never farther than the raw tiles inside it, so its
            // level is at least as fine)
            used_farc = L.far_coarse_any && !pair;
            if (used_farc) { a.far_mode = 1; a.dfar = L.dfar; }
            if ((rc_code = launch_gather<1, false>(a, tmap, L.raw, N, st)) != DRM_OK) return rc_code;
            a.far_mode = 0;
            if (pair) {
                f.part = PART_FAR; f.route_mask = ROUTE_SPEC_RAW;
                if ((rc_code = launch_gather<1, false>(f, tmap, L.far, N, st)) != DRM_OK) return rc_code;
            }
            if (used_farc) {
                GatherArgs c2 = g;
                c2.src = L.coarse2_map; c2.Hm = L.Hc2; c2.Wm = L.Wc2; c2.slab = L.slab_farc;
                c2.dth_cell = (float)(COARSE2 * M_PI / He); c2.dph_cell = (float)(COARSE2 * 2.0 * M_PI / We);
                fill_plan(c2, L.farc_raw);
                set_levels(c2, S, hierarchy, S);
                c2.part = PART_ALL; c2.route_mask = ROUTE_SPEC_RAW;
                c2.far_mode = 2; c2.dfar = L.dfar;
                c2.raw_tt = L.raw.tt; c2.raw_ttiles_x = L.raw.ttiles_x; c2.raw_Hm = He; c2.raw_Wm = We;
                c2.raw_dth = a.dth_cell; c2.raw_dph = a.dph_cell;
                CUtensorMap tmapc;
                memset(&tmapc, 0, sizeof(tmapc));
                c2.use_tma = ((L.Wc2 * COARSE_FLOATS * 4) % 16 == 0);
                if (c2.use_tma && (rc_code = make_tensor_map(&tmapc, L.coarse2_map, B, L.Hc2, L.Wc2, COARSE_FLOATS, FARC_TT)) != DRM_OK) return rc_code;
                c2.far_factor = COARSE2; c2.dfar_hi = L.far_coarse4 ? L.dfar4 : 0.f;
                if ((rc_code = launch_gather<1, true>(c2, tmapc, L.farc_raw, N, st)) != DRM_OK) return rc_code;
                if (L.far_coarse4) {
                    GatherArgs c4 = c2;
                    c4.src = L.coarse_map; c4.Hm = L.Hc; c4.Wm = L.Wc; c4.slab = L.slab_farc4;
                    c4.dth_cell = (float)(COARSE * M_PI / He); c4.dph_cell = (float)(COARSE * 2.0 * M_PI / We);
                    fill_plan(c4, L.farc4_raw);
                    c4.far_factor = COARSE; c4.dfar = L.dfar4; c4.dfar_hi = 0.f;
                    CUtensorMap tmap4;
                    memset(&tmap4, 0, sizeof(tmap4));
                    c4.use_tma = ((L.Wc * COARSE_FLOATS * 4) % 16 == 0);
                    if (c4.use_tma && (rc_code = make_tensor_map(&tmap4, L.coarse_map, B, L.Hc, L.Wc, COARSE_FLOATS, FARC_TT)) != DRM_OK) return rc_code;
                    if ((rc_code = launch_gather<1, true>(c4, tmap4, L.farc4_raw, N, st)) != DRM_OK) return rc_code;
                }
            }
        }
        if (!L.coarse_diffuse_ok) {  // otherwise no render is routed to BOTH_RAW
            a.part = pair ? PART_NEAR : PART_ALL;
            a.route_mask = ROUTE_BOTH_RAW;
            if ((rc_code = launch_gather<3, false>(a, tmap, L.raw, N, st)) != DRM_OK) return rc_code;
            if (pair) {
                f.part = PART_FAR; f.route_mask = ROUTE_BOTH_RAW;
                if ((rc_code = launch_gather<3, false>(f, tmap, L.far, N, st)) != DRM_OK) return rc_code;
            }
        }
        used_far = pair;
    }
    // ---- launches on the coarse map: diffuse lobe of the renders above, both lobes of the very rough renders --------
    if (L.coarse_enabled) {
        GatherArgs a = g;
        a.src = L.coarse_map; a.Hm = L.Hc; a.Wm = L.Wc;
        a.dth_cell = (float)(COARSE * M_PI / He); a.dph_cell = (float)(COARSE * 2.0 * M_PI / We);
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        a.use_tma = ((L.Wc * COARSE_FLOATS * 4) % 16 == 0);
        if (a.use_tma && (rc_code = make_tensor_map(&tmap, L.coarse_map, B, L.Hc, L.Wc, COARSE_FLOATS, TT)) != DRM_OK) return rc_code;
        if (L.coarse_diffuse_ok) {
            a.slab = L.slab_diff; a.route_mask = ROUTE_DIFF_COARSE;
            const RenderPlan& dp = L.diff_corr ? L.diff1 : L.diff;
            fill_plan(a, dp);
            set_levels(a, dp.S, false);
            if ((rc_code = launch_gather<2, true>(a, tmap, dp, N, st)) != DRM_OK) return rc_code;
            if (L.diff_corr) {
                // + (2x2 lattice - centre sample) on the 16x16-texel map
                GatherArgs c16 = g;
                c16.src = L.coarse16_map; c16.Hm = L.Hc16; c16.Wm = L.Wc16;
                c16.dth_cell = (float)(4 * COARSE * M_PI / He); c16.dph_cell = (float)(4 * COARSE * 2.0 * M_PI / We);
                c16.route_mask = ROUTE_DIFF_COARSE;
                CUtensorMap tmap16;
                memset(&tmap16, 0, sizeof(tmap16));
                c16.use_tma = ((L.Wc16 * COARSE_FLOATS * 4) % 16 == 0);
                if (c16.use_tma && (rc_code = make_tensor_map(&tmap16, L.coarse16_map, B, L.Hc16, L.Wc16, COARSE_FLOATS, TT)) != DRM_OK) return rc_code;
                c16.slab = L.slab_corr_hi;
                fill_plan(c16, L.corr_hi);
                set_levels(c16, 2, false);
                if ((rc_code = launch_gather<2, true>(c16, tmap16, L.corr_hi, N, st)) != DRM_OK) return rc_code;
                c16.slab = L.slab_corr_lo;
                fill_plan(c16, L.corr_lo);
                set_levels(c16, 1, false);
                if ((rc_code = launch_gather<2, true>(c16, tmap16, L.corr_lo, N, st)) != DRM_OK) return rc_code;
            }
        }
        a.slab = L.slab_coarse; a.route_mask = ROUTE_BOTH_COARSE;
        fill_plan(a, L.coarse);
        set_levels(a, S, hierarchy, S);
        if ((rc_code = launch_gather<3, true>(a, tmap, L.coarse, N, st)) != DRM_OK) return rc_code;
        // the 2x2 coarsening: both lobes of the moderately rough renders
        GatherArgs a2 = g;
        a2.src = L.coarse2_map; a2.Hm = L.Hc2; a2.Wm = L.Wc2;
        a2.dth_cell = (float)(COARSE2 * M_PI / He); a2.dph_cell = (float)(COARSE2 * 2.0 * M_PI / We);
        CUtensorMap tmap2;
        memset(&tmap2, 0, sizeof(tmap2));
        a2.use_tma = ((L.Wc2 * COARSE_FLOATS * 4) % 16 == 0);
        if (a2.use_tma && (rc_code = make_tensor_map(&tmap2, L.coarse2_map, B, L.Hc2, L.Wc2, COARSE_FLOATS, TT)) != DRM_OK) return rc_code;
        a2.slab = L.slab_coarse2; a2.route_mask = ROUTE_BOTH_COARSE2;
        fill_plan(a2, L.coarse2);
        set_levels(a2, S, hierarchy, S);
        if ((rc_code = launch_gather<3, true>(a2, tmap2, L.coarse2, N, st)) != DRM_OK) return rc_code;
    }
    {
        SlabSet set{};
        set.s[0] = SlabDesc{L.slab_raw, L.raw.splits, (near_mode ? 0 : ROUTE_SPEC_RAW) | ROUTE_BOTH_RAW, 1.f};
        set.s[1] = SlabDesc{L.coarse_diffuse_ok ? L.slab_diff : nullptr, L.diff_corr ? L.diff1.splits : L.diff.splits,
                            ROUTE_DIFF_COARSE, 1.f};
        set.s[2] = SlabDesc{L.coarse_enabled ? L.slab_coarse : nullptr, L.coarse.splits, ROUTE_BOTH_COARSE, 1.f};
        set.s[3] = SlabDesc{(used_far || near_mode) ? L.slab_far : nullptr,
                            (near_mode && used_farc) ? L.mid.splits : L.far.splits,
                            ((used_far || near_mode) ? ROUTE_SPEC_RAW : 0) | (used_far ? ROUTE_BOTH_RAW : 0), 1.f};
        set.s[4] = SlabDesc{L.coarse_enabled ? L.slab_coarse2 : nullptr, L.coarse2.splits, ROUTE_BOTH_COARSE2, 1.f};
        set.s[5] = SlabDesc{near_mode ? L.slab_near : nullptr, 1, ROUTE_SPEC_RAW, 1.f};
        set.s[6] = SlabDesc{used_farc ? L.slab_farc : nullptr, near_mode ? L.farc.splits : L.farc_raw.splits, ROUTE_SPEC_RAW, 1.f};
        set.s[7] = SlabDesc{L.diff_corr ? L.slab_corr_hi : nullptr, L.corr_hi.splits, ROUTE_DIFF_COARSE, 1.f};
        set.s[8] = SlabDesc{L.diff_corr ? L.slab_corr_lo : nullptr, L.corr_lo.splits, ROUTE_DIFF_COARSE, -1.f};
        set.s[9] = SlabDesc{(used_farc && L.far_coarse4) ? L.slab_farc4 : nullptr, near_mode ? L.farc4.splits : L.farc4_raw.splits,
                            ROUTE_SPEC_RAW, 1.f};
        const size_t total = (size_t)N * res * res * 3;
        render_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(set, L.rc, out, N, res, channel_first);
        count_launches(1);
    }
    DRM_CHECK_CUDA(cudaGetLastError());
    return DRM_OK;
}
