// Single-level evaluation of the refmap render: every (sub-normal, texel) pair of the defining sum
//
//   out[k, i, j, c] = sum_{a,b < S} w_a w_b  sum_texels  f_c(d_t; v_k, n(theta_i,a, phi_j,b); z_k) E[t, c] dOmega_t
//
// This is the round-1 tile kernel reduced to its core and kept as the VALIDATION path (drm_render_refmaps_flat): whole
// images at sizes the fp64 oracle cannot reach, in the same fp32 arithmetic as the product path but without any of its
// approximations.  One CTA owns a tile of refmap cells of one render (1024 sub-normal slots: S x S Gauss-Legendre
// nodes per cell) and streams tt x tt-texel envmap tiles through shared memory with TMA (cp.async.bulk.tensor, double
// buffered, mbarrier completion); a cooperative transform turns each raw tile into pixel-independent records (half
// vector, |v + d|, Fresnel-weighted radiance * solid angle); every thread gathers the records for its 4 slots.  The
// texel tiles are split over blockIdx.z; render_flat_combine_kernel sums the partial slabs in fixed order.
// ~1e4 times more work than drm_render_refmaps (render_tree.cu); bound by the FP32 / MUFU pipes.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace drm {

static constexpr int FLAT_THREADS = 256;
static constexpr int FLAT_SUBS = 4;
static constexpr int FLAT_SLOTS = FLAT_THREADS * FLAT_SUBS;
static constexpr int FLAT_TT = 32;  // largest tile edge in texels
static constexpr int FLAT_TILE = FLAT_TT * FLAT_TT;

struct FlatConst {  // per render
    float vhat[3], left[3], upp[3];
    float m, rough, alpha2, inv_a2m1, one_m_a2, eta;
    float base[3], cdiff[3];
    int env, pad;
};

struct FlatArgs {
    const float* env;
    const FlatConst* rc;
    const float *sin_t, *cos_t, *sin_p, *cos_p;
    float* slab;  // [splits][N][res*res][3]
    int B, He, We, N, res, S, G;
    int tile_e, tiles_x;      // cells per CTA edge, CTAs per refmap row
    int tt, ttiles_x, ntiles, splits, use_tma;
    float domega_k, cell;
    float gl_x[16], gl_w[16];
};

__global__ void flat_tables_kernel(float* sin_t, float* cos_t, float* sin_p, float* cos_p, int He, int We) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < He) {
        const double t = (i + 0.5) * (M_PI / He);
        sin_t[i] = (float)sin(t);
        cos_t[i] = (float)cos(t);
    }
    if (i < We) {
        const double p = (i + 0.5) * (2.0 * M_PI / We);
        sin_p[i] = (float)sin(p);
        cos_p[i] = (float)cos(p);
    }
}

// clip z to [0,1] (mitsuba3_utils.py:239,242), BSDF constants, camera frame of look_at(v, 0, +Y) (:235-236)
__global__ void flat_setup_kernel(const float* __restrict__ z6, const float* __restrict__ view3, const uint8_t* __restrict__ flip,
                                  const int32_t* __restrict__ env_index, int N, int B, float alpha_min, FlatConst* __restrict__ rc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    FlatConst c;
    float z[6];
    for (int i = 0; i < 6; ++i) z[i] = fminf(fmaxf(z6[6 * k + i], 0.f), 1.f);
    c.m = z[0];
    c.base[0] = z[1]; c.base[1] = z[2]; c.base[2] = z[3];
    c.rough = z[4];
    const float alpha = fmaxf(z[4] * z[4], alpha_min);
    c.alpha2 = alpha * alpha;
    c.inv_a2m1 = 1.f / c.alpha2 - 1.f;
    c.one_m_a2 = 1.f - c.alpha2;
    c.eta = 2.f / (1.f - sqrtf(0.08f * z[5])) - 1.f;
    for (int i = 0; i < 3; ++i) c.cdiff[i] = (1.f - c.m) * c.base[i] * (float)M_1_PI;
    float vx = view3[3 * k], vy = view3[3 * k + 1], vz = view3[3 * k + 2];
    const float inv = rsqrtf(vx * vx + vy * vy + vz * vz);
    vx *= inv; vy *= inv; vz *= inv;
    c.vhat[0] = vx; c.vhat[1] = vy; c.vhat[2] = vz;
    const float fx = -vx, fy = -vy, fz = -vz;
    float lx = fz, ly = 0.f, lz = -fx;  // up x forward, up = (0,1,0)
    const float linv = rsqrtf(lx * lx + lz * lz);
    lx *= linv; lz *= linv;
    c.upp[0] = fy * lz - fz * ly;
    c.upp[1] = fz * lx - fx * lz;
    c.upp[2] = fx * ly - fy * lx;
    const float sgn = (flip && flip[k]) ? -1.f : 1.f;
    c.left[0] = sgn * lx; c.left[1] = sgn * ly; c.left[2] = sgn * lz;
    const int e = env_index ? env_index[k] : k;
    c.env = min(max(e, 0), B - 1);
    c.pad = 0;
    rc[k] = c;
}

__device__ __forceinline__ float flat_fresnel(float cos_i, float eta) {
    const float eta_ti = 1.f / eta;
    const float ct2 = 1.f - eta_ti * eta_ti * (1.f - cos_i * cos_i);
    if (ct2 <= 0.f) return 1.f;
    const float ct = sqrtf(ct2);
    const float a_s = (cos_i - eta * ct) / (cos_i + eta * ct);
    const float a_p = (ct - eta * cos_i) / (ct + eta * cos_i);
    return 0.5f * (a_s * a_s + a_p * a_p);
}

__global__ void __launch_bounds__(FLAT_THREADS, 2)
render_flat_kernel(const __grid_constant__ CUtensorMap tmap, const FlatArgs g) {
    constexpr int RAW_FLOATS = FLAT_TILE * 3;
    const int tt = g.tt, tile_texels = tt * tt, row_floats = tt * 3;
    extern __shared__ __align__(1024) unsigned char flat_smem[];
    float* raw0 = reinterpret_cast<float*>(flat_smem);
    float4* rec = reinterpret_cast<float4*>(flat_smem + 2 * RAW_FLOATS * sizeof(float));
    uint64_t* bars = reinterpret_cast<uint64_t*>(flat_smem + 2 * RAW_FLOATS * sizeof(float) + FLAT_TILE * 48);

    const int tid = threadIdx.x;
    const int k = blockIdx.y;
    const FlatConst rc = g.rc[k];
    const int pty = blockIdx.x / g.tiles_x, ptx = blockIdx.x - pty * g.tiles_x;
    const int pi0 = pty * g.tile_e, pj0 = ptx * g.tile_e;
    const int G = g.G, npix = g.tile_e * g.tile_e;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        if (g.use_tma) tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    // my 4 slots: slot q -> cell q / G, Gauss-Legendre node q % G
    float nx[FLAT_SUBS], ny[FLAT_SUBS], nz[FLAT_SUBS], nv[FLAT_SUBS], Fi[FLAT_SUBS], mult[FLAT_SUBS], wq[FLAT_SUBS];
#pragma unroll
    for (int r = 0; r < FLAT_SUBS; ++r) {
        const int q = r * FLAT_THREADS + tid;
        const int pl = q / G, node = q - pl * G;
        const int li = pl / g.tile_e, lj = pl - li * g.tile_e;
        const int i = pi0 + li, j = pj0 + lj;
        const int a = node / g.S, b = node - a * g.S;
        const bool active = pl < npix && i < g.res && j < g.res;
        const float th = ((float)i + 0.5f + 0.5f * g.gl_x[a]) * g.cell;
        const float ph = ((float)j + 0.5f + 0.5f * g.gl_x[b]) * g.cell;
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
        wq[r] = active ? g.gl_w[a] * g.gl_w[b] : 0.f;
        // F D G1(n.v) G1(n.d) / (4 n.v) = F x / (q^2 (x + sq)) / (pi alpha^2 (n.v + sqrt((n.v)^2 (1-a^2) + a^2)))
        const float g1 = lz + sqrtf(lz * lz * rc.one_m_a2 + rc.alpha2);
        mult[r] = lz > 0.f ? wq[r] / (3.14159265358979f * rc.alpha2 * g1) : 0.f;
        if (!(lz > 0.f)) wq[r] = 0.f;
    }

    const int tbeg = blockIdx.z, tstep = g.splits;
    const int nmine = (g.ntiles - tbeg + tstep - 1) / tstep;
    auto issue = [&](int e, int stage) {
        const int tile = tbeg + e * tstep;
        const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
        mbar_arrive_expect_tx(&bars[stage], tile_texels * 3 * sizeof(float));
        tma_load_3d(raw0 + stage * RAW_FLOATS, &tmap, &bars[stage], tx * row_floats, ty * tt, rc.env);
    };
    if (g.use_tma && tid == 0) {
        if (0 < nmine) issue(0, 0);
        if (1 < nmine) issue(1, 1);
    }

    float tot[FLAT_SUBS][6];
#pragma unroll
    for (int r = 0; r < FLAT_SUBS; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) tot[r][c] = 0.f;

    for (int it = 0; it < nmine; ++it) {
        const int stage = it & 1;
        float* raw = raw0 + stage * RAW_FLOATS;
        const int tile = tbeg + it * tstep;
        const int ty = tile / g.ttiles_x, tx = tile - ty * g.ttiles_x;
        if (g.use_tma) {
            mbar_wait(&bars[stage], (it >> 1) & 1);
        } else {
            // plain-load staging for maps whose row pitch is not a multiple of 16 bytes
            const int map_row = g.We * 3;
            const float* src = g.env + (size_t)rc.env * g.He * map_row;
            for (int e = tid; e < tile_texels * 3; e += FLAT_THREADS) {
                const int lr = e / row_floats, lc = e - lr * row_floats;
                const int r = ty * tt + lr, c = tx * row_floats + lc;
                raw[e] = (r < g.He && c < map_row) ? src[(size_t)r * map_row + c] : 0.f;
            }
            __syncthreads();
        }
        // transform: raw tile -> records {h, |v+d|, Rr, E dOmega F_c, E dOmega}
        for (int t = tid; t < tile_texels; t += FLAT_THREADS) {
            const int lr = t / tt, lc = t - lr * tt;
            const bool inside = ty * tt + lr < g.He && tx * tt + lc < g.We;  // TMA fills out-of-range texels with zeros
            const int r = min(ty * tt + lr, g.He - 1), c = min(tx * tt + lc, g.We - 1);
            const float st = g.sin_t[r], ct = g.cos_t[r], sp = g.sin_p[c], cp = g.cos_p[c];
            const float dx = st * sp, dy = ct, dz = -st * cp;
            const float dom = inside ? g.domega_k * st : 0.f;
            const float er = raw[lr * row_floats + lc * 3 + 0] * dom;
            const float eg = raw[lr * row_floats + lc * 3 + 1] * dom;
            const float eb = raw[lr * row_floats + lc * 3 + 2] * dom;
            // |v + d|^2 from its components: 2 + 2 v.d cancels at grazing reflection (d ~ -v) and cost 3 % at the limb cells
            const float sx = rc.vhat[0] + dx, sy = rc.vhat[1] + dy, sz = rc.vhat[2] + dz;
            const float len2 = fmaxf(sx * sx + sy * sy + sz * sz, 1e-12f);
            const float inv_len = rsqrtf(len2);
            const float len = len2 * inv_len;
            const float vh = 0.5f * len;
            const float Fd = flat_fresnel(vh, rc.eta);
            const float mm = fminf(fmaxf(1.f - vh, 0.f), 1.f);
            const float sw = (mm * mm) * (mm * mm) * mm;
            const float fr = (1.f - rc.m) * Fd + rc.m * (rc.base[0] + (1.f - rc.base[0]) * sw);
            const float fg = (1.f - rc.m) * Fd + rc.m * (rc.base[1] + (1.f - rc.base[1]) * sw);
            const float fb = (1.f - rc.m) * Fd + rc.m * (rc.base[2] + (1.f - rc.base[2]) * sw);
            rec[t * 3 + 0] = make_float4(sx * inv_len, sy * inv_len, sz * inv_len, len);
            rec[t * 3 + 1] = make_float4(2.f * rc.rough * vh * vh, er * fr, eg * fg, eb * fb);
            rec[t * 3 + 2] = make_float4(er, eg, eb, 0.f);
        }
        __syncthreads();  // records ready, raw[stage] free
        if (g.use_tma && tid == 0 && it + 2 < nmine) {
            fence_proxy_async();
            issue(it + 2, stage);
        }
        float acc[FLAT_SUBS][6];
#pragma unroll
        for (int r = 0; r < FLAT_SUBS; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) acc[r][c] = 0.f;
#pragma unroll 2
        for (int t = 0; t < tile_texels; ++t) {
            const float4 h = rec[t * 3 + 0];
            const float4 s = rec[t * 3 + 1];
            const float4 d4 = rec[t * 3 + 2];
#pragma unroll
            for (int r = 0; r < FLAT_SUBS; ++r) {
                const float ex = nx[r] - h.x, ey = ny[r] - h.y, ez = nz[r] - h.z;
                const float u2 = ex * ex + ey * ey + ez * ez;  // 2 (1 - n.h), no cancellation
                const float nh = 1.f - 0.5f * u2;
                const float xc = fmaxf(h.w * nh - nv[r], 0.f);  // n.d = |v+d| n.h - n.v; below the horizon: weight 0
                const float sin2 = u2 * (1.f - 0.25f * u2);
                const float q = 1.f + sin2 * rc.inv_a2m1;
                const float sq = fast_sqrt(xc * xc * rc.one_m_a2 + rc.alpha2);
                const float ws = xc * fast_rcp(q * q * (xc + sq));
                acc[r][0] += ws * s.y; acc[r][1] += ws * s.z; acc[r][2] += ws * s.w;
                const float mm = 1.f - xc;
                const float m2 = mm * mm;
                const float Fo = m2 * m2 * mm;
                const float Rr = s.x;
                const float inner = (-0.5f + 0.25f * Fi[r]) + Rr * ((1.f - Fi[r]) + Rr * Fi[r]);
                const float wd = xc * (((1.f - 0.5f * Fi[r]) + Rr * Fi[r]) + Fo * inner);
                acc[r][3] += wd * d4.x; acc[r][4] += wd * d4.y; acc[r][5] += wd * d4.z;
            }
        }
        // two-level summation (per tile, then total) keeps the fp32 error near 1e-6
#pragma unroll
        for (int r = 0; r < FLAT_SUBS; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) tot[r][c] += mult[r] * acc[r][c];
#pragma unroll
            for (int c = 3; c < 6; ++c) tot[r][c] += wq[r] * acc[r][c];
        }
        __syncthreads();  // records free
    }

    // per-cell reduction over its S^2 slots in fixed order
    float* resbuf = reinterpret_cast<float*>(rec);  // [FLAT_SLOTS][3]
#pragma unroll
    for (int r = 0; r < FLAT_SUBS; ++r) {
        const int q = r * FLAT_THREADS + tid;
#pragma unroll
        for (int c = 0; c < 3; ++c) resbuf[q * 3 + c] = tot[r][c] + rc.cdiff[c] * tot[r][3 + c];
    }
    __syncthreads();
    for (int o = tid; o < npix * 3; o += FLAT_THREADS) {
        const int pl = o / 3, c = o - pl * 3;
        const int li = pl / g.tile_e, lj = pl - li * g.tile_e;
        const int i = pi0 + li, j = pj0 + lj;
        if (i >= g.res || j >= g.res) continue;
        float v = 0.f;
        for (int s2 = 0; s2 < G; ++s2) v += resbuf[(pl * G + s2) * 3 + c];
        const size_t pix = (size_t)i * g.res + j;
        g.slab[(((size_t)blockIdx.z * g.N + k) * g.res * g.res + pix) * 3 + c] = v;
    }
}

__global__ void render_flat_combine_kernel(const float* __restrict__ slab, int splits, float* __restrict__ out, int N, int res,
                                           int channel_first) {
    const size_t total = (size_t)N * res * res * 3;
    const size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int c = (int)(o % 3);
    const size_t pix = (o / 3) % ((size_t)res * res);
    const size_t k = o / 3 / ((size_t)res * res);
    float v = 0.f;
    for (int sp = 0; sp < splits; ++sp) v += slab[(size_t)sp * total + o];
    out[channel_first ? (k * 3 + c) * res * res + pix : o] = v;
}

static void flat_gauss_legendre(int S, float* x, float* w) {
    for (int i = 0; i < S; ++i) {
        double z = cos(M_PI * (i + 0.75) / (S + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < S; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
            }
            pp = S * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-15) break;
        }
        x[S - 1 - i] = (float)z;
        w[S - 1 - i] = (float)(1.0 / ((1.0 - z * z) * pp * pp));  // = w_i / 2: the weights sum to 1
    }
}

struct FlatLayout {
    int tile_e, tiles_x, tt, ttiles_x, ntiles, splits;
    FlatConst* rc;
    float *sin_t, *cos_t, *sin_p, *cos_p, *slab;
};

static size_t flat_layout(FlatLayout& L, void* ws, int N, int B, int He, int We, int res, int S) {
    (void)B;
    int e = (int)floor(sqrt((double)(FLAT_SLOTS / (S * S))));
    if (e < 1) e = 1;
    if (e > res) e = res;
    L.tile_e = e;
    L.tiles_x = (res + e - 1) / e;
    L.tt = He >= 1600 ? 32 : He >= 400 ? 16 : 8;
    L.ttiles_x = (We + L.tt - 1) / L.tt;
    L.ntiles = L.ttiles_x * ((He + L.tt - 1) / L.tt);
    const long ctas = (long)L.tiles_x * L.tiles_x * N;
    long s = (148L * 2 * 4 + ctas - 1) / ctas;  // a few waves of two resident CTAs per SM
    if (s > L.ntiles / 4) s = L.ntiles / 4;
    if (s < 1) s = 1;
    L.splits = (int)s;
    Carver c(ws);
    L.rc = c.take<FlatConst>(N);
    L.sin_t = c.take<float>(He);
    L.cos_t = c.take<float>(He);
    L.sin_p = c.take<float>(We);
    L.cos_p = c.take<float>(We);
    L.slab = c.take<float>((size_t)N * res * res * 3 * L.splits);
    return c.used();
}

}  // namespace drm

using namespace drm;

extern "C" size_t drm_render_flat_workspace_bytes(int N, int B, int He, int We, int res, int S) {
    if (N <= 0 || B <= 0 || He <= 0 || We <= 0 || res <= 0 || S < 1 || S > 16) return 0;
    FlatLayout L;
    return flat_layout(L, nullptr, N, B, He, We, res, S);
}

extern "C" int drm_render_refmaps_flat(const float* env, int B, int He, int We, const int32_t* env_index, const float* z6,
                                       const float* view3, const uint8_t* flip, int N, int res, int S, float alpha_min,
                                       int channel_first, float* out, void* workspace, size_t workspace_bytes,
                                       void* cuda_stream) {
    DRM_REQUIRE(env && z6 && view3 && out, "render_flat: null pointer");
    DRM_REQUIRE(N > 0 && B > 0 && He > 0 && We > 0 && res > 0, "render_flat: N=%d B=%d He=%d We=%d res=%d must be positive", N, B, He, We, res);
    DRM_REQUIRE(S >= 1 && S <= 16, "render_flat: footprint_S=%d not in 1..16", S);
    DRM_REQUIRE(res <= 4096 && N <= 65535, "render_flat: res=%d / N=%d too large", res, N);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    FlatLayout L;
    const size_t need = flat_layout(L, workspace, N, B, He, We, res, S);
    if (!workspace || workspace_bytes < need) {
        set_error("render_flat: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return DRM_EWORKSPACE;
    }
    if (!(alpha_min > 0.f)) alpha_min = fmaxf(1e-3f, (float)(1.25 * M_PI / He));
    FlatArgs g;
    memset(&g, 0, sizeof(g));
    g.env = env; g.rc = L.rc; g.sin_t = L.sin_t; g.cos_t = L.cos_t; g.sin_p = L.sin_p; g.cos_p = L.cos_p; g.slab = L.slab;
    g.B = B; g.He = He; g.We = We; g.N = N; g.res = res; g.S = S; g.G = S * S;
    g.tile_e = L.tile_e; g.tiles_x = L.tiles_x; g.tt = L.tt; g.ttiles_x = L.ttiles_x; g.ntiles = L.ntiles; g.splits = L.splits;
    g.domega_k = (float)((2.0 * M_PI / We) * (M_PI / He));
    g.cell = (float)(M_PI / res);
    flat_gauss_legendre(S, g.gl_x, g.gl_w);
    const int tb = 128;
    flat_tables_kernel<<<(max(He, We) + tb - 1) / tb, tb, 0, st>>>(L.sin_t, L.cos_t, L.sin_p, L.cos_p, He, We);
    flat_setup_kernel<<<(N + tb - 1) / tb, tb, 0, st>>>(z6, view3, flip, env_index, N, B, alpha_min, L.rc);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    g.use_tma = ((We * 12) % 16 == 0) && ((reinterpret_cast<uintptr_t>(env) & 15) == 0);
    if (g.use_tma) {
        PFN_encodeTiled enc = get_encode_tiled();
        if (!enc) {
            set_error("render_flat: cuTensorMapEncodeTiled entry point unavailable");
            return DRM_ECUDA;
        }
        const cuuint64_t row = (cuuint64_t)We * 3;
        cuuint64_t dims[3] = {row, (cuuint64_t)He, (cuuint64_t)B};
        cuuint64_t strides[2] = {row * 4, row * 4 * (cuuint64_t)He};
        cuuint32_t box[3] = {(cuuint32_t)(L.tt * 3), (cuuint32_t)L.tt, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(env), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("render_flat: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return DRM_ECUDA;
        }
    }
    const size_t smem = 2 * (size_t)FLAT_TILE * 3 * sizeof(float) + (size_t)FLAT_TILE * 48 + 16;
    DRM_CHECK_CUDA(cudaFuncSetAttribute(render_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    render_flat_kernel<<<dim3(L.tiles_x * L.tiles_x, N, L.splits), FLAT_THREADS, smem, st>>>(tmap, g);
    const size_t total = (size_t)N * res * res * 3;
    render_flat_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L.slab, L.splits, out, N, res, channel_first);
    count_launches(4);
    DRM_CHECK_CUDA(cudaGetLastError());
    return DRM_OK;
}
