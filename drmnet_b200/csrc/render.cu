// K1: reflectance-map forward render as a tiled gather-reduce over the environment map (sm_100a).
//
// Replaces MitsubaRefMapRenderer.rendering (reference utils/mitsuba3_utils.py:411-430 -> :365-409 -> :217-246): the
// Mitsuba scene {unit sphere, `principled` BSDF, lat-long envmap emitter, normal-indexed orthographic sensor, `direct`
// integrator, box filter} evaluated as its deterministic limit
//
//   out[k, i, j, c] = sum_{a,b < S} w_a w_b  sum_texels  f_c(d_t; v_k, n(theta_i,a, phi_j,b); z_k) E[t, c] dOmega_t
//
// with S x S Gauss-Legendre sub-normals per refmap cell (the box pixel filter) and the texel-centre quadrature of the
// emitter.  One CTA owns a tile of sub-normals of one render and streams 32x32-texel envmap tiles through shared
// memory with TMA (cp.async.bulk.tensor, double buffered, mbarrier completion); a cooperative transform turns each raw
// tile into per-texel records (half vector, Fresnel-weighted radiance * solid angle, retro-reflection factor) that are
// pixel independent, then every thread gathers all 1024 records for its 4 sub-normals from shared memory (broadcast
// LDS.128).  The kernel is bound by the FP32/MUFU pipes (about 34 instructions per (sub-normal, texel) pair), not by
// HBM: each envmap byte is reused by every sub-normal of the render out of L2.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace drm {

static constexpr int GATHER_THREADS = 256;
static constexpr int SUBS_PER_THREAD = 4;
static constexpr int SLOTS = GATHER_THREADS * SUBS_PER_THREAD;  // sub-normal slots per CTA
static constexpr int TT = 32;                                   // texel tile edge
static constexpr int TILE_TEXELS = TT * TT;
static constexpr int RAW_FLOATS = TT * TT * 3;
static constexpr int REC_FLOATS = 12;
static constexpr int MAX_LEVELS = 5;      // footprint lattices 1,2,4,8,16 per axis
static constexpr int MAX_LIST = 2048;     // texel tiles one CTA can schedule (plan splits larger maps)

struct RenderConst {  // per render
    float vhat[3], left[3], upp[3];  // camera frame of look_at(v, 0, +Y); `left` carries the flip sign
    float m, rough, alpha2, inv_a2m1, one_m_a2, eta;
    float base[3], cdiff[3];
    float thr[MAX_LEVELS];  // half-vector-space distance beyond which footprint level k is accurate enough
    int env, has_diffuse;
};

struct GatherArgs {
    const float* env;
    const RenderConst* rc;
    const float *sin_t, *cos_t, *sin_p, *cos_p;
    float* out;
    float* partial;
    int B, He, We, N, res, S;
    int tile_w, tile_h, tiles_x, tiles_y;
    int ttiles_x, ttiles_y, splits;
    int channel_first, use_tma, cull;
    int nlev;                 // number of footprint levels used by this launch
    int lev_S[MAX_LEVELS];    // lattice size per axis of level k (ascending; the last one is S)
    int lev_tidx[MAX_LEVELS]; // log2(lev_S[k]): index into RenderConst::thr
    float domega_k, cell;
    float gl_x[MAX_LEVELS][16], gl_w[MAX_LEVELS][16];
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

// clip z to [0,1] (mitsuba3_utils.py:239,242), derive the BSDF constants and the camera frame (:235-236)
__global__ void render_setup_kernel(const float* __restrict__ z6, const float* __restrict__ view3,
                                    const uint8_t* __restrict__ flip, const int32_t* __restrict__ env_index, int N,
                                    int B, float alpha_min, float cell, float level_scale,
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
        c.thr[0] = 21.0f * sqrtf(ca);                                  // 1 x 1
        c.thr[1] = 7.5f * powf(cell, 2.f / 3.f) * powf(alpha, 1.f / 3.f);  // 2 x 2
        c.thr[2] = 2.4f * powf(cell, 0.8f) * powf(alpha, 0.2f);        // 4 x 4
        c.thr[3] = 1.2f * powf(cell, 8.f / 9.f) * powf(alpha, 1.f / 9.f);  // 8 x 8
        c.thr[4] = 0.f;                                                // 16 x 16
        for (int i = 0; i < MAX_LEVELS - 1; ++i) c.thr[i] = level_scale * fmaxf(c.thr[i], 6.f * alpha);
    }
    for (int i = 0; i < 3; ++i) c.cdiff[i] = (1.f - c.m) * c.base[i] * (float)M_1_PI;
    c.has_diffuse = (c.cdiff[0] > 0.f || c.cdiff[1] > 0.f || c.cdiff[2] > 0.f) ? 1 : 0;
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

__device__ __forceinline__ float fresnel_dielectric(float cos_i, float eta) {
    float eta_ti = 1.f / eta;
    float ct2 = 1.f - eta_ti * eta_ti * (1.f - cos_i * cos_i);
    if (ct2 <= 0.f) return 1.f;
    float ct = sqrtf(ct2);
    float a_s = (cos_i - eta * ct) / (cos_i + eta * ct);
    float a_p = (ct - eta * cos_i) / (ct + eta * cos_i);
    return 0.5f * (a_s * a_s + a_p * a_p);
}

// Classify one texel tile for a CTA whose normals lie in the cone (axis a, radius beta):
//   0      no normal of the cone sees any texel of the tile (n.d <= 0 everywhere): skipped
//   1+k    footprint level k (0 = 1x1 lattice ... nlev-1 = full S x S lattice) chosen from the distance, in
//          half-vector space, between the cone of normals and the tile's half vectors h = normalize(v + d)
__device__ __forceinline__ int classify_tile(const GatherArgs& g, const float* __restrict__ vhat,
                                             const float* __restrict__ thr, int tile, float ax, float ay, float az,
                                             float beta) {
    const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
    const int r0 = ty * TT, r1 = min(r0 + TT, g.He), c0 = tx * TT, c1 = min(c0 + TT, g.We);
    const float dth = 0.5f * (r1 - r0) * (3.14159265f / g.He), dph = 0.5f * (c1 - c0) * (6.2831853f / g.We);
    const float thc = 0.5f * (r0 + r1) * (3.14159265f / g.He), phc = 0.5f * (c0 + c1) * (6.2831853f / g.We);
    float st, ct, sp, cp;
    sincosf(thc, &st, &ct);
    sincosf(phc, &sp, &cp);
    // angular radius of the tile: meridian move + parallel move (the arc on the parallel bounds the great-circle one)
    const float gamma = dth + dph * fminf(1.f, st + dth);
    const float dx = st * sp, dy = ct, dz = -st * cp;
    if (g.cull) {
        const float spread = beta + gamma + 0.01f;
        if (spread < 1.5607963f && ax * dx + ay * dy + az * dz <= -sinf(spread)) return 0;
    }
    if (g.nlev == 1) return 1;
    const float hx = vhat[0] + dx, hy = vhat[1] + dy, hz = vhat[2] + dz;
    const float len = sqrtf(hx * hx + hy * hy + hz * hz);
    if (len - gamma < 0.05f) return g.nlev;  // d ~ -v: the d -> h map is singular, stay on the finest lattice
    const float gamma_h = gamma / (len - gamma);  // |dh| <= |dd| / |v + d|
    const float cosang = fminf(fmaxf((ax * hx + ay * hy + az * hz) / len, -1.f), 1.f);
    const float dist = acosf(cosang) - beta - gamma_h;
    for (int k = 0; k < g.nlev - 1; ++k)
        if (dist >= thr[g.lev_tidx[k]]) return 1 + k;
    return g.nlev;
}

template <bool HAS_DIFFUSE>
__global__ void __launch_bounds__(GATHER_THREADS, 2)
render_gather_kernel(const __grid_constant__ CUtensorMap tmap, const GatherArgs g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* raw0 = reinterpret_cast<float*>(smem_raw);
    float4* rec = reinterpret_cast<float4*>(smem_raw + 2 * RAW_FLOATS * sizeof(float));
    unsigned char* tail = smem_raw + 2 * RAW_FLOATS * sizeof(float) + TILE_TEXELS * REC_FLOATS * sizeof(float);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);                 // 2 mbarriers
    int* scan_ws = reinterpret_cast<int*>(tail + 16);                     // 8 warp sums + level starts
    uint16_t* list = reinterpret_cast<uint16_t*>(tail + 16 + 64);         // [MAX_LIST] tiles ordered by level
    uint8_t* lvl = reinterpret_cast<uint8_t*>(tail + 16 + 64 + MAX_LIST * 2);  // [MAX_LIST]

    const int tid = threadIdx.x;
    const int k = blockIdx.y;
    const RenderConst rc = g.rc[k];
    if (HAS_DIFFUSE != (rc.has_diffuse != 0)) return;  // the other instantiation serves this render

    const int ptile = blockIdx.x;
    const int pty = ptile / g.tiles_x, ptx = ptile - pty * g.tiles_x;
    const int pi0 = pty * g.tile_h, pj0 = ptx * g.tile_w;
    const int S2 = g.S * g.S;
    const int npix = g.tile_w * g.tile_h;

    // ---- cone of this CTA's normals (cell corners included) ---------------------------------------------------------
    float ax, ay, az, beta;
    {
        const int i1 = min(pi0 + g.tile_h, g.res), j1 = min(pj0 + g.tile_w, g.res);
        const float thc = 0.5f * (pi0 + i1) * g.cell, phc = 0.5f * (pj0 + j1) * g.cell;
        float st, ct, sp, cp;
        sincosf(thc, &st, &ct);
        sincosf(phc, &sp, &cp);
        const float lx = st * cp, lz = st * sp;
        ax = lx * rc.left[0] + ct * rc.upp[0] + lz * rc.vhat[0];
        ay = lx * rc.left[1] + ct * rc.upp[1] + lz * rc.vhat[1];
        az = lx * rc.left[2] + ct * rc.upp[2] + lz * rc.vhat[2];
        const float dth = 0.5f * (i1 - pi0) * g.cell, dph = 0.5f * (j1 - pj0) * g.cell;
        beta = dth + dph * fminf(1.f, st + dth);
    }

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        if (g.use_tma) tma_prefetch_desc(&tmap);
    }

    // ---- schedule: classify my texel tiles, then order them by footprint level (coarse first, tile order inside) ----
    const int ntiles = g.ttiles_x * g.ttiles_y;
    const int per = (ntiles + g.splits - 1) / g.splits;
    const int tbeg = blockIdx.z * per, tend = min(tbeg + per, ntiles);
    const int nmine = tend - tbeg;
    for (int e = tid; e < nmine; e += GATHER_THREADS)
        lvl[e] = (uint8_t)classify_tile(g, g.rc[k].vhat, g.rc[k].thr, tbeg + e, ax, ay, az, beta);
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
        const int tile = tbeg + list[e];
        const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
        mbar_arrive_expect_tx(&bars[stage], RAW_FLOATS * sizeof(float));
        tma_load_3d(raw0 + stage * RAW_FLOATS, &tmap, &bars[stage], tx * TT * 3, ty * TT, rc.env);
    };
    if (g.use_tma && tid == 0) {
        if (0 < nlist) issue(0, 0);
        if (1 < nlist) issue(1, 1);
    }

    // ---- per-level state of my 4 slots: lattice node, texel subset, normals -----------------------------------------
    // slot q = r*256 + tid -> pixel q / S^2, sub = q % S^2.  At a level with an Sk x Sk lattice the S^2 slots of a
    // pixel form Sk^2 groups of gk = (S/Sk)^2 slots: the group evaluates one lattice node, its gk slots split the tile's
    // texels (t = u, u+gk, ...).  S^2 divides 256, so the 4 slots of a thread share node and subset.
    float nx[SUBS_PER_THREAD], ny[SUBS_PER_THREAD], nz[SUBS_PER_THREAD], nv[SUBS_PER_THREAD];
    float Fi[SUBS_PER_THREAD], mult[SUBS_PER_THREAD], wq[SUBS_PER_THREAD];
    int cur_level = 0, level_end = 0, gk = 1, u0 = 0;
    const bool hier = (GATHER_THREADS % S2) == 0;

    for (int it = 0; it < nlist; ++it) {
        if (it >= level_end) {
            // next non-empty level
            do { ++cur_level; level_end = scan_ws[8 + cur_level]; } while (it >= level_end);
            const int Sk = g.lev_S[cur_level - 1];
            const int ratio = g.S / Sk;
            gk = hier ? ratio * ratio : 1;
#pragma unroll
            for (int r = 0; r < SUBS_PER_THREAD; ++r) {
                const int q = r * GATHER_THREADS + tid;
                const int pl = q / S2, sub = q - pl * S2;
                const int li = pl / g.tile_w, lj = pl - li * g.tile_w;
                const int i = pi0 + li, j = pj0 + lj;
                const int node = sub / gk;
                if (r == 0) u0 = sub - node * gk;
                const int a = node / Sk, b = node - a * Sk;
                const bool active = pl < npix && i < g.res && j < g.res;
                const float th = ((float)i + 0.5f + 0.5f * g.gl_x[cur_level - 1][a]) * g.cell;
                const float ph = ((float)j + 0.5f + 0.5f * g.gl_x[cur_level - 1][b]) * g.cell;
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
            }
        }
        const int stage = it & 1;
        float* raw = raw0 + stage * RAW_FLOATS;
        const int tile = tbeg + list[it];
        const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
        if (g.use_tma) {
            mbar_wait(&bars[stage], (it >> 1) & 1);
        } else {
            // plain-load fallback for maps whose row pitch is not a multiple of 16 bytes
            const float* src = g.env + (size_t)rc.env * g.He * g.We * 3;
            for (int e = tid; e < RAW_FLOATS; e += GATHER_THREADS) {
                const int lr = e / (TT * 3), lc3 = e - lr * (TT * 3);
                const int r = ty * TT + lr, c3 = tx * TT * 3 + lc3;
                raw[e] = (r < g.He && c3 < g.We * 3) ? src[(size_t)r * g.We * 3 + c3] : 0.f;
            }
            __syncthreads();
        }

        // ---- transform: raw RGB -> pixel-independent per-texel records ------------------------------------------
#pragma unroll
        for (int u = 0; u < TILE_TEXELS / GATHER_THREADS; ++u) {
            const int t = u * GATHER_THREADS + tid;
            const int lr = t / TT, lc = t - lr * TT;
            const int r = min(ty * TT + lr, g.He - 1), c = min(tx * TT + lc, g.We - 1);
            const float st = g.sin_t[r], ct = g.cos_t[r], sp = g.sin_p[c], cp = g.cos_p[c];
            const float dx = st * sp, dy = ct, dz = -st * cp;
            const float dom = g.domega_k * st;
            const float vd = rc.vhat[0] * dx + rc.vhat[1] * dy + rc.vhat[2] * dz;
            const float len2 = fmaxf(2.f + 2.f * vd, 1e-12f);
            const float inv_len = rsqrtf(len2);
            const float len = len2 * inv_len;
            const float vh = 0.5f * len;
            const float Fd = fresnel_dielectric(vh, rc.eta);
            const float mm = fminf(fmaxf(1.f - vh, 0.f), 1.f);
            const float sw = (mm * mm) * (mm * mm) * mm;
            const float er = raw[lr * TT * 3 + lc * 3 + 0] * dom;
            const float eg = raw[lr * TT * 3 + lc * 3 + 1] * dom;
            const float eb = raw[lr * TT * 3 + lc * 3 + 2] * dom;
            const float fr = (1.f - rc.m) * Fd + rc.m * (rc.base[0] + (1.f - rc.base[0]) * sw);
            const float fg = (1.f - rc.m) * Fd + rc.m * (rc.base[1] + (1.f - rc.base[1]) * sw);
            const float fb = (1.f - rc.m) * Fd + rc.m * (rc.base[2] + (1.f - rc.base[2]) * sw);
            rec[t * 3 + 0] = make_float4((rc.vhat[0] + dx) * inv_len, (rc.vhat[1] + dy) * inv_len,
                                         (rc.vhat[2] + dz) * inv_len, len);
            rec[t * 3 + 1] = make_float4(2.f * rc.rough * vh * vh, er * fr, eg * fg, eb * fb);
            rec[t * 3 + 2] = make_float4(er, eg, eb, 0.f);
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
        for (int t = u0; t < TILE_TEXELS; t += gk) {
            const float4 h = rec[t * 3 + 0];
            const float4 s = rec[t * 3 + 1];
            float4 d4;
            if (HAS_DIFFUSE) d4 = rec[t * 3 + 2];
#pragma unroll
            for (int r = 0; r < SUBS_PER_THREAD; ++r) {
                const float ex = nx[r] - h.x, ey = ny[r] - h.y, ez = nz[r] - h.z;
                const float u2 = ex * ex + ey * ey + ez * ez;         // 2 (1 - n.h), no cancellation
                const float nh = 1.f - 0.5f * u2;
                const float x = h.w * nh - nv[r];                     // n.d = |v+d| n.h - n.v
                const float sin2 = u2 * (1.f - 0.25f * u2);           // 1 - (n.h)^2
                const float q = 1.f + sin2 * rc.inv_a2m1;             // cos^2 + sin^2 / alpha^2
                const float xc = fmaxf(x, 0.f);                       // below the horizon: weight 0, denominator > 0
                const float sq = fast_sqrt(xc * xc * rc.one_m_a2 + rc.alpha2);
                const float ws = xc * fast_rcp(q * q * (xc + sq));    // D G1(n.d) up to per-slot constants
                acc[r][0] += ws * s.y;
                acc[r][1] += ws * s.z;
                acc[r][2] += ws * s.w;
                if (HAS_DIFFUSE) {
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
#pragma unroll
            for (int c = 0; c < 3; ++c) tot[r][c] += mult[r] * acc[r][c];
            if (HAS_DIFFUSE) {
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
            float v = tot[r][c];
            if (HAS_DIFFUSE) v += rc.cdiff[c] * tot[r][3 + c];
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
        for (int s2 = 0; s2 < S2; ++s2) v += resbuf[(pl * S2 + s2) * 3 + c];
        const size_t pix = (size_t)i * g.res + j;
        if (g.splits == 1) {
            const size_t idx = g.channel_first ? ((size_t)k * 3 + c) * g.res * g.res + pix
                                               : ((size_t)k * g.res * g.res + pix) * 3 + c;
            g.out[idx] = v;
        } else {
            g.partial[(((size_t)blockIdx.z * g.N + k) * g.res * g.res + pix) * 3 + c] = v;
        }
    }
}

__global__ void render_reduce_splits_kernel(const float* __restrict__ partial, float* __restrict__ out, int N, int res,
                                            int splits, int channel_first) {
    const size_t total = (size_t)N * res * res * 3;
    size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (o >= total) return;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += partial[(size_t)s * total + o];  // fixed order: deterministic
    const int c = (int)(o % 3);
    const size_t pix = (o / 3) % ((size_t)res * res);
    const size_t k = o / 3 / ((size_t)res * res);
    const size_t idx = channel_first ? (k * 3 + c) * res * res + pix : o;
    out[idx] = v;
}

struct RenderPlan {
    int tile_w, tile_h, tiles_x, tiles_y, ttiles_x, ttiles_y, splits;
};

static RenderPlan make_plan(int N, int He, int We, int res, int S) {
    RenderPlan p;
    int px = SLOTS / (S * S);
    int e = (int)floor(sqrt((double)px));
    if (e < 1) e = 1;
    if (e > res) e = res;
    p.tile_w = p.tile_h = e;
    p.tiles_x = (res + e - 1) / e;
    p.tiles_y = (res + e - 1) / e;
    p.ttiles_x = (We + TT - 1) / TT;
    p.ttiles_y = (He + TT - 1) / TT;
    const long ctas = (long)p.tiles_x * p.tiles_y * N;
    const long want = 148L * 2 * 2;  // two waves of two resident CTAs per SM
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

struct RenderWs {
    RenderConst* rc;
    float *sin_t, *cos_t, *sin_p, *cos_p, *partial;
};

static size_t render_carve(RenderWs& w, void* ws, int N, int He, int We, int res, const RenderPlan& p) {
    Carver c(ws);
    w.rc = c.take<RenderConst>(N);
    w.sin_t = c.take<float>(He);
    w.cos_t = c.take<float>(He);
    w.sin_p = c.take<float>(We);
    w.cos_p = c.take<float>(We);
    w.partial = c.take<float>(p.splits > 1 ? (size_t)p.splits * N * res * res * 3 : 1);
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

}  // namespace drm

using namespace drm;

extern "C" size_t drm_render_workspace_bytes(int N, int B, int He, int We, int res, int S) {
    if (N <= 0 || B <= 0 || He <= 0 || We <= 0 || res <= 0 || S < 1 || S > 16) return 0;
    RenderWs w;
    return render_carve(w, nullptr, N, He, We, res, make_plan(N, He, We, res, S));
}

extern "C" int drm_render_refmaps(const float* env, int B, int He, int We, const int32_t* env_index, const float* z6,
                                  const float* view3, const uint8_t* flip, int N, int res, int S, float alpha_min,
                                  int channel_first, float* out, void* workspace, size_t workspace_bytes,
                                  void* cuda_stream) {
    DRM_REQUIRE(env && z6 && view3 && out, "render: null pointer");
    DRM_REQUIRE(N > 0 && B > 0 && He > 0 && We > 0 && res > 0, "render: N=%d B=%d He=%d We=%d res=%d must be positive", N, B, He, We, res);
    DRM_REQUIRE(S >= 1 && S <= 16, "render: footprint_S=%d not in 1..16", S);
    DRM_REQUIRE(res <= 4096, "render: res=%d too large", res);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const RenderPlan p = make_plan(N, He, We, res, S);
    RenderWs w;
    const size_t need = render_carve(w, workspace, N, He, We, res, p);
    if (!workspace || workspace_bytes < need) {
        set_error("render: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return DRM_EWORKSPACE;
    }
    if (!(alpha_min > 0.f)) alpha_min = fmaxf(1e-3f, (float)(1.25 * M_PI / He));

    GatherArgs g{};
    g.env = env; g.rc = w.rc; g.sin_t = w.sin_t; g.cos_t = w.cos_t; g.sin_p = w.sin_p; g.cos_p = w.cos_p;
    g.out = out; g.partial = w.partial;
    g.B = B; g.He = He; g.We = We; g.N = N; g.res = res; g.S = S;
    g.tile_w = p.tile_w; g.tile_h = p.tile_h; g.tiles_x = p.tiles_x; g.tiles_y = p.tiles_y;
    g.ttiles_x = p.ttiles_x; g.ttiles_y = p.ttiles_y; g.splits = p.splits;
    g.channel_first = channel_first;
    g.cull = 1;
    g.domega_k = (float)((2.0 * M_PI / We) * (M_PI / He));
    g.cell = (float)(M_PI / res);
    // footprint levels: power-of-two lattices below S when S is one of 2,4,8,16 (S^2 then divides the 256 threads);
    // any other S runs as a single level
    g.nlev = 0;
    const bool pow2 = (S == 2 || S == 4 || S == 8 || S == 16);
    const char* lv = getenv("DRM_RENDER_LEVELS");  // "0" disables the hierarchy (debugging / validation)
    const bool hierarchy = pow2 && !(lv && lv[0] == '0');
    if (hierarchy)
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
    float level_scale = 0.3f;  // thresholds of render_setup_kernel are conservative; 0.3 measured (scripts/levels_probe.py)
    if (const char* ls = getenv("DRM_RENDER_LEVEL_SCALE")) level_scale = (float)atof(ls);

    // TMA descriptor over env viewed as [B][He][3*We] fp32; box = 32 rows x 96 floats of one map
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    g.use_tma = ((We * 12) % 16 == 0) && ((reinterpret_cast<uintptr_t>(env) & 15) == 0);
    if (g.use_tma) {
        PFN_encodeTiled enc = get_encode_tiled();
        if (!enc) {
            set_error("render: cuTensorMapEncodeTiled entry point unavailable");
            return DRM_ECUDA;
        }
        cuuint64_t dims[3] = {(cuuint64_t)We * 3, (cuuint64_t)He, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)We * 12, (cuuint64_t)We * 12 * (cuuint64_t)He};
        cuuint32_t box[3] = {TT * 3, TT, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(env), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("render: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return DRM_ECUDA;
        }
    }

    const int tb = 128;
    render_tables_kernel<<<(max(He, We) + tb - 1) / tb, tb, 0, st>>>(w.sin_t, w.cos_t, w.sin_p, w.cos_p, He, We);
    render_setup_kernel<<<(N + tb - 1) / tb, tb, 0, st>>>(z6, view3, flip, env_index, N, B, alpha_min, g.cell, level_scale, w.rc);

    const size_t smem = 2 * RAW_FLOATS * sizeof(float) + TILE_TEXELS * REC_FLOATS * sizeof(float) + 16 + 64 + MAX_LIST * 3;
    DRM_CHECK_CUDA(cudaFuncSetAttribute(render_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DRM_CHECK_CUDA(cudaFuncSetAttribute(render_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(p.tiles_x * p.tiles_y, N, p.splits);
    // both instantiations are launched; each CTA exits at once when its render belongs to the other one
    render_gather_kernel<false><<<grid, GATHER_THREADS, smem, st>>>(tmap, g);
    render_gather_kernel<true><<<grid, GATHER_THREADS, smem, st>>>(tmap, g);
    if (p.splits > 1) {
        const size_t total = (size_t)N * res * res * 3;
        render_reduce_splits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w.partial, out, N, res, p.splits, channel_first);
    }
    DRM_CHECK_CUDA(cudaGetLastError());
    count_launches(4 + (p.splits > 1 ? 1 : 0));
    return DRM_OK;
}
