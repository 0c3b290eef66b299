/*
 * Development replica (fp64, CPU) of the hierarchical render of drmnet_b200/csrc/render_tree.cu: per-render source pyramid
 * with half-vector moments, dual traversal by pixel-block passes, second-order (covariance) cell evaluation.
 * Used to tune the acceptance constants against the fp64 oracle before spending GPU time.  Not part of the product.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXLEV 12

typedef struct {
    double mu[3], S[6], w[3], rh, vh2;
    double mc[3][3];                     /* per channel: sum_t w_c,t (h_t - mu): the channels' own centroids */   /* spec record: half-vector mean, covariance (xx,yy,zz,xy,xz,yz), F-weighted energy */
    double md[3], Sd[6], e[3], rd;
    double dc[3][3];                     /* per channel: sum_t e_c,t (d_t - md) */       /* diffuse record: direction mean, covariance, energy */
} Rec;

typedef struct {
    int He, We, L;
    int H[MAXLEV], W[MAXLEV];
    Rec* lev[MAXLEV]; /* lev[0] = texels */
} Pyr;

typedef struct {
    double vhat[3], left[3], upp[3];
    double m, rough, alpha, alpha2, inv_a2m1, one_m_a2, eta, base[3], cdiff[3];
    double thr[5];
    double kappa, kappa_d, hz, rcap, hand;
    int pk;      /* log2 S */
    int res;
    double cell;
    const double* glx[5];
    const double* glw[5];
    int terms;
    int pixcov, full2, chan;
} RC;

typedef struct { long visits, pairs0, pairs1, handed; } Stats;

static double fresnel_dielectric(double cos_i, double eta) {
    double eta_ti = 1.0 / eta;
    double ct2 = 1.0 - eta_ti * eta_ti * (1.0 - cos_i * cos_i);
    if (ct2 <= 0.0) return 1.0;
    double ct = sqrt(ct2);
    double a_s = (cos_i - eta * ct) / (cos_i + eta * ct);
    double a_p = (ct - eta * cos_i) / (ct + eta * cos_i);
    return 0.5 * (a_s * a_s + a_p * a_p);
}

static void camera_frame(const double* view, double* vhat, double* left, double* upp) {
    double len = sqrt(view[0] * view[0] + view[1] * view[1] + view[2] * view[2]);
    for (int k = 0; k < 3; ++k) vhat[k] = view[k] / len;
    double fwd[3] = {-vhat[0], -vhat[1], -vhat[2]};
    double l[3] = {fwd[2], 0.0, -fwd[0]};
    double ll = sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (int k = 0; k < 3; ++k) left[k] = l[k] / ll;
    upp[0] = fwd[1] * left[2] - fwd[2] * left[1];
    upp[1] = fwd[2] * left[0] - fwd[0] * left[2];
    upp[2] = fwd[0] * left[1] - fwd[1] * left[0];
}

/* merge n child records into a parent */
static void merge(Rec* p, const Rec** ch, int n) {
    double om = 0, m1[3] = {0, 0, 0}, m2[6] = {0, 0, 0, 0, 0, 0}, w[3] = {0, 0, 0};
    double od = 0, d1[3] = {0, 0, 0}, d2[6] = {0, 0, 0, 0, 0, 0}, e[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const Rec* c = ch[i];
        double o = c->w[0] + c->w[1] + c->w[2];
        if (o > 0) {
            om += o;
            for (int k = 0; k < 3; ++k) { m1[k] += o * c->mu[k]; w[k] += c->w[k]; }
            m2[0] += o * (c->S[0] + c->mu[0] * c->mu[0]);
            m2[1] += o * (c->S[1] + c->mu[1] * c->mu[1]);
            m2[2] += o * (c->S[2] + c->mu[2] * c->mu[2]);
            m2[3] += o * (c->S[3] + c->mu[0] * c->mu[1]);
            m2[4] += o * (c->S[4] + c->mu[0] * c->mu[2]);
            m2[5] += o * (c->S[5] + c->mu[1] * c->mu[2]);
        }
        double q = c->e[0] + c->e[1] + c->e[2];
        if (q > 0) {
            od += q;
            for (int k = 0; k < 3; ++k) { d1[k] += q * c->md[k]; e[k] += c->e[k]; }
            d2[0] += q * (c->Sd[0] + c->md[0] * c->md[0]);
            d2[1] += q * (c->Sd[1] + c->md[1] * c->md[1]);
            d2[2] += q * (c->Sd[2] + c->md[2] * c->md[2]);
            d2[3] += q * (c->Sd[3] + c->md[0] * c->md[1]);
            d2[4] += q * (c->Sd[4] + c->md[0] * c->md[2]);
            d2[5] += q * (c->Sd[5] + c->md[1] * c->md[2]);
        }
    }
    memset(p, 0, sizeof(*p));
    if (om > 0) {
        for (int k = 0; k < 3; ++k) { p->mu[k] = m1[k] / om; p->w[k] = w[k]; }
        p->S[0] = m2[0] / om - p->mu[0] * p->mu[0];
        p->S[1] = m2[1] / om - p->mu[1] * p->mu[1];
        p->S[2] = m2[2] / om - p->mu[2] * p->mu[2];
        p->S[3] = m2[3] / om - p->mu[0] * p->mu[1];
        p->S[4] = m2[4] / om - p->mu[0] * p->mu[2];
        p->S[5] = m2[5] / om - p->mu[1] * p->mu[2];
        double r = 0;
        for (int i = 0; i < n; ++i) {
            const Rec* c = ch[i];
            if (c->w[0] + c->w[1] + c->w[2] <= 0) continue;
            double dx = c->mu[0] - p->mu[0], dy = c->mu[1] - p->mu[1], dz = c->mu[2] - p->mu[2];
            double d = sqrt(dx * dx + dy * dy + dz * dz) + c->rh;
            if (d > r) r = d;
        }
        p->rh = r;
        for (int i = 0; i < n; ++i) {
            const Rec* c = ch[i];
            if (c->w[0] + c->w[1] + c->w[2] <= 0) continue;
            for (int cc = 0; cc < 3; ++cc)
                for (int k = 0; k < 3; ++k) p->mc[cc][k] += c->mc[cc][k] + c->w[cc] * (c->mu[k] - p->mu[k]);
        }
    }
    if (od > 0) {
        for (int k = 0; k < 3; ++k) { p->md[k] = d1[k] / od; p->e[k] = e[k]; }
        p->Sd[0] = d2[0] / od - p->md[0] * p->md[0];
        p->Sd[1] = d2[1] / od - p->md[1] * p->md[1];
        p->Sd[2] = d2[2] / od - p->md[2] * p->md[2];
        p->Sd[3] = d2[3] / od - p->md[0] * p->md[1];
        p->Sd[4] = d2[4] / od - p->md[0] * p->md[2];
        p->Sd[5] = d2[5] / od - p->md[1] * p->md[2];
        double r = 0;
        for (int i = 0; i < n; ++i) {
            const Rec* c = ch[i];
            if (c->e[0] + c->e[1] + c->e[2] <= 0) continue;
            double dx = c->md[0] - p->md[0], dy = c->md[1] - p->md[1], dz = c->md[2] - p->md[2];
            double d = sqrt(dx * dx + dy * dy + dz * dz) + c->rd;
            if (d > r) r = d;
        }
        p->rd = r;
        for (int i = 0; i < n; ++i) {
            const Rec* c = ch[i];
            if (c->e[0] + c->e[1] + c->e[2] <= 0) continue;
            for (int cc = 0; cc < 3; ++cc)
                for (int k = 0; k < 3; ++k) p->dc[cc][k] += c->dc[cc][k] + c->e[cc] * (c->md[k] - p->md[k]);
        }
    }
}

static void build_pyramid(Pyr* P, const float* env, int He, int We, const RC* rc) {
    P->He = He; P->We = We;
    P->H[0] = He; P->W[0] = We;
    P->lev[0] = (Rec*)calloc((size_t)He * We, sizeof(Rec));
    const double dom_k = (2.0 * M_PI / We) * (M_PI / He);
#pragma omp parallel for
    for (int r = 0; r < He; ++r) {
        double t = (r + 0.5) * (M_PI / He), st = sin(t), ct = cos(t);
        for (int c = 0; c < We; ++c) {
            double p = (c + 0.5) * (2.0 * M_PI / We);
            double d[3] = {st * sin(p), ct, -st * cos(p)};
            Rec* R = &P->lev[0][(size_t)r * We + c];
            const float* e = env + ((size_t)r * We + c) * 3;
            double dom = dom_k * st;
            double vd = rc->vhat[0] * d[0] + rc->vhat[1] * d[1] + rc->vhat[2] * d[2];
            double len2 = 2.0 + 2.0 * vd;
            if (len2 < 1e-12) len2 = 1e-12;
            double inv_len = 1.0 / sqrt(len2);
            double vh = 0.5 * len2 * inv_len;
            for (int k = 0; k < 3; ++k) R->mu[k] = (rc->vhat[k] + d[k]) * inv_len;
            double Fd = fresnel_dielectric(vh, rc->eta);
            double mm = 1.0 - vh; if (mm < 0) mm = 0; if (mm > 1) mm = 1;
            double sw = mm * mm * mm * mm * mm;
            for (int k = 0; k < 3; ++k) {
                double F = (1.0 - rc->m) * Fd + rc->m * (rc->base[k] + (1.0 - rc->base[k]) * sw);
                R->w[k] = e[k] * dom * F;
                R->e[k] = e[k] * dom;
                R->md[k] = d[k];
            }
            R->vh2 = 2.0 * vh;
        }
    }
    int l = 0;
    while (P->H[l] * P->W[l] > 256 && l + 1 < MAXLEV) {
        int Hn = (P->H[l] + 1) / 2, Wn = (P->W[l] + 1) / 2;
        P->H[l + 1] = Hn; P->W[l + 1] = Wn;
        P->lev[l + 1] = (Rec*)calloc((size_t)Hn * Wn, sizeof(Rec));
#pragma omp parallel for
        for (int r = 0; r < Hn; ++r)
            for (int c = 0; c < Wn; ++c) {
                const Rec* ch[4]; int n = 0;
                for (int dr = 0; dr < 2; ++dr)
                    for (int dc = 0; dc < 2; ++dc) {
                        int rr = 2 * r + dr, cc = 2 * c + dc;
                        if (rr < P->H[l] && cc < P->W[l]) ch[n++] = &P->lev[l][(size_t)rr * P->W[l] + cc];
                    }
                Rec* R = &P->lev[l + 1][(size_t)r * Wn + c];
                merge(R, ch, n);
                double m2 = R->mu[0] * R->mu[0] + R->mu[1] * R->mu[1] + R->mu[2] * R->mu[2];
                double vm = rc->vhat[0] * R->mu[0] + rc->vhat[1] * R->mu[1] + rc->vhat[2] * R->mu[2];
                R->vh2 = m2 > 1e-30 ? 2.0 * vm / m2 : 0.0;
            }
        ++l;
    }
    P->L = l;
}

static double view_term(const RC* rc, double lz) {
    if (lz <= 0) return 0;
    double g1 = lz + sqrt(lz * lz * rc->one_m_a2 + rc->alpha2);
    return 1.0 / (M_PI * rc->alpha2 * g1);
}

typedef struct { double n[3], nv, mult, wq, Fi; double q[6]; double P[6]; int pc; } Node;

/* node (I,J) of the lattice 2^p: normal, weights (view-term averaged when p < pk) */
static void make_node(const RC* rc, int p, int I, int J, int flip, Node* nd) {
    int Sk = 1 << p, S = 1 << rc->pk;
    int i = I >> p, a = I & (Sk - 1), j = J >> p, b = J & (Sk - 1);
    double sa = 0, sb = 0, meanv = -1;
    if (p < rc->pk) {
        int m = S / Sk;
        const double* fx = rc->glx[rc->pk];
        const double* fw = rc->glw[rc->pk];
        double num = 0, den = 0, va = 0, vb = 0, ua = 0, ub = 0;
        for (int ia = 0; ia < m; ++ia) {
            double xa = fx[a * m + ia], wa = fw[a * m + ia];
            double st = sin((i + 0.5 + 0.5 * xa) * rc->cell);
            for (int ib = 0; ib < m; ++ib) {
                double xb = fx[b * m + ib];
                double sp = sin((j + 0.5 + 0.5 * xb) * rc->cell);
                double w = wa * fw[b * m + ib];
                double wv = w * view_term(rc, st * sp);
                num += wv; den += w; va += wv * xa; vb += wv * xb; ua += w * xa; ub += w * xb;
            }
        }
        if (num > 0) { meanv = num / den; sa = va / num - ua / den; sb = vb / num - ub / den; }
        else meanv = 0;
    }
    double th = (i + 0.5 + 0.5 * (rc->glx[p][a] + sa)) * rc->cell;
    double ph = (j + 0.5 + 0.5 * (rc->glx[p][b] + sb)) * rc->cell;
    double st = sin(th), ct = cos(th), sp = sin(ph), cp = cos(ph);
    double lx = (flip ? -1.0 : 1.0) * st * cp, lz = st * sp;
    for (int k = 0; k < 3; ++k) nd->n[k] = lx * rc->left[k] + ct * rc->upp[k] + lz * rc->vhat[k];
    nd->nv = lz;
    nd->wq = rc->glw[p][a] * rc->glw[p][b];
    nd->mult = nd->wq * (meanv >= 0 ? meanv : view_term(rc, lz));
    double mm = 1.0 - lz; if (mm < 0) mm = 0; if (mm > 1) mm = 1;
    nd->Fi = mm * mm * mm * mm * mm;
    nd->q[0] = nd->n[0] * nd->n[0]; nd->q[1] = nd->n[1] * nd->n[1]; nd->q[2] = nd->n[2] * nd->n[2];
    nd->q[3] = 2 * nd->n[0] * nd->n[1]; nd->q[4] = 2 * nd->n[0] * nd->n[2]; nd->q[5] = 2 * nd->n[1] * nd->n[2];
    /* pixel-side covariance of the node's sub-cell (uniform box in theta, phi of half-widths cell/2/Sk) */
    memset(nd->P, 0, sizeof(nd->P));
    nd->pc = rc->pixcov && p == 0 && rc->pk > 0;
    if (rc->pixcov && p == 0 && rc->pk > 0) {
        double hw = rc->cell / Sk;            /* full width of the sub-cell (approximate for GL sub-regions) */
        double var = hw * hw / 12.0;
        double et[3], ep[3];
        double lxt = (flip ? -1.0 : 1.0) * ct * cp, lzt = ct * sp;      /* d n / d theta */
        double lxp = (flip ? -1.0 : 1.0) * (-st * sp), lzp = st * cp;   /* d n / d phi */
        for (int k = 0; k < 3; ++k) {
            et[k] = lxt * rc->left[k] - st * rc->upp[k] + lzt * rc->vhat[k];
            ep[k] = lxp * rc->left[k] + lzp * rc->vhat[k];
        }
        nd->P[0] = var * (et[0] * et[0] + ep[0] * ep[0]);
        nd->P[1] = var * (et[1] * et[1] + ep[1] * ep[1]);
        nd->P[2] = var * (et[2] * et[2] + ep[2] * ep[2]);
        nd->P[3] = var * (et[0] * et[1] + ep[0] * ep[1]);
        nd->P[4] = var * (et[0] * et[2] + ep[0] * ep[2]);
        nd->P[5] = var * (et[1] * et[2] + ep[1] * ep[2]);
    }
}

static inline void eval_pair(const RC* rc, const Node* nd, const Rec* R, int level, double* acc) {
    if (rc->terms & 1) {
        double ex = nd->n[0] - R->mu[0], ey = nd->n[1] - R->mu[1], ez = nd->n[2] - R->mu[2];
        double trS = R->S[0] + R->S[1] + R->S[2];
        double u = ex * ex + ey * ey + ez * ez + trS;     /* mean of 2 - 2 n.h over the cell (exact) */
        double nmu = 1.0 - 0.5 * u;                        /* n . mu */
        const double* S = R->S; const double* n = nd->n; const double* v = rc->vhat;
        double nSn = 0, nSv = 0, vSv = 0, vmu = 0;
        int second = level > 0 || nd->pc;
        if (second) {
            nSn = nd->q[0] * S[0] + nd->q[1] * S[1] + nd->q[2] * S[2] + nd->q[3] * S[3] + nd->q[4] * S[4] + nd->q[5] * S[5];
            if (rc->full2) {
                nSv = n[0] * v[0] * S[0] + n[1] * v[1] * S[1] + n[2] * v[2] * S[2] + (n[0] * v[1] + n[1] * v[0]) * S[3] +
                      (n[0] * v[2] + n[2] * v[0]) * S[4] + (n[1] * v[2] + n[2] * v[1]) * S[5];
                vSv = v[0] * v[0] * S[0] + v[1] * v[1] * S[1] + v[2] * v[2] * S[2] +
                      2 * (v[0] * v[1] * S[3] + v[0] * v[2] * S[4] + v[1] * v[2] * S[5]);
            }
        }
        vmu = v[0] * R->mu[0] + v[1] * R->mu[1] + v[2] * R->mu[2];
        double x;
        if (rc->full2 && level > 0) x = 2 * vmu * nmu + 2 * nSv - nd->nv;   /* mean of n.d over the cell (exact) */
        else x = R->vh2 * nmu - nd->nv;
        /* a cell that straddles the horizon of the normal (n.d = 0): the lobe is cut there and ramps up over ~alpha, so the
           clamp is applied to the distribution of n.d inside the cell (uniform of the cell's variance), not to its mean */
        double varx = 0; int straddle = 0;
        if (rc->full2 && level > 0) {
            varx = 4.0 * (nmu * nmu * vSv + 2.0 * nmu * vmu * nSv + vmu * vmu * nSn);
            double w = sqrt(3.0 * fmax(varx, 0.0));
            if (x < w) {
                straddle = 1;
                x = x > -w ? (x + w) * (x + w) / (4.0 * w) : 0.0;
            }
        }
        if (x > 0 && nd->nv > 0) {
            double sin2 = u * (1.0 - 0.25 * u);
            double q = 1.0 + sin2 * rc->inv_a2m1;
            double rq = 1.0 / q;
            double D = rq * rq;
            double sq = sqrt(x * x * rc->one_m_a2 + rc->alpha2);
            double g = x / (x + sq);
            double K = D * g;
            if (second) {
                double pS = nSn;
                if (nd->pc) {
                    const double* m = R->mu;
                    pS += m[0] * m[0] * nd->P[0] + m[1] * m[1] * nd->P[1] + m[2] * m[2] * nd->P[2] +
                          2 * (m[0] * m[1] * nd->P[3] + m[0] * m[2] * nd->P[4] + m[1] * m[2] * nd->P[5]);
                }
                double qu = rc->inv_a2m1 * (1.0 - 0.5 * u);
                double a = qu * rq;
                double Duu = D * (6.0 * a * a + rc->inv_a2m1 * rq);     /* D_uu */
                K += 2.0 * pS * Duu * g;
                if (rc->full2 && level > 0) {
                    double Du = -2.0 * D * a;
                    double xs = x + sq;
                    double gx = rc->alpha2 / (sq * xs * xs);
                    double sp = x * rc->one_m_a2 / sq;
                    double gxx = -rc->alpha2 * (sp / (sq * sq * xs * xs) + 2.0 * (1.0 + sp) / (sq * xs * xs * xs));
                    double T2 = -8.0 * Du * gx * (nmu * nSv + vmu * nSn);
                    double T3 = straddle ? 0.0 : D * gxx * varx;
                    K += 0.5 * (T2 + T3);
                }
            }
            double ws = nd->mult * K;
            acc[0] += ws * R->w[0]; acc[1] += ws * R->w[1]; acc[2] += ws * R->w[2];
            if (rc->chan && level > 0) {
                double qu = rc->inv_a2m1 * (1.0 - 0.5 * u);
                double Du = -2.0 * D * qu * rq;
                double xs = x + sq;
                double gx = rc->alpha2 / (sq * xs * xs);
                double cn = -2.0 * Du * g + (rc->chan > 1 ? 2.0 * D * gx * vmu : 0.0);
                double cv = rc->chan > 1 ? 2.0 * D * gx * nmu : 0.0;
                for (int c = 0; c < 3; ++c) {
                    double nm = n[0] * R->mc[c][0] + n[1] * R->mc[c][1] + n[2] * R->mc[c][2];
                    double vm = v[0] * R->mc[c][0] + v[1] * R->mc[c][1] + v[2] * R->mc[c][2];
                    acc[c] += nd->mult * (cn * nm + cv * vm);
                }
            }
        }
    }
    if (rc->terms & 2) {
        /* diffuse lobe, second order in the direction spread: x = n.d, y = 1 + v.d are linear in d */
        const double* d = R->md;
        const double* S = R->Sd;
        const double* n = nd->n; const double* v = rc->vhat;
        double x = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
        double nSn = 0; int straddle = 0;
        if (level > 0) {
            nSn = n[0] * n[0] * S[0] + n[1] * n[1] * S[1] + n[2] * n[2] * S[2] +
                  2 * (n[0] * n[1] * S[3] + n[0] * n[2] * S[4] + n[1] * n[2] * S[5]);
            double w = sqrt(3.0 * fmax(nSn, 0.0));
            if (x < w) { straddle = 1; x = x > -w ? (x + w) * (x + w) / (4.0 * w) : 0.0; }
        }
        if (x > 0 && nd->nv > 0) {
            double y = 1.0 + v[0] * d[0] + v[1] * d[1] + v[2] * d[2];
            double Fi = nd->Fi, r = rc->rough;
            /* K = x [ A + B y + Fo (C + D y + E y^2) ],  Fo = (1-x)^5 */
            double A = 1.0 - 0.5 * Fi, B = r * Fi, C = -0.5 * (1.0 - 0.5 * Fi), Dd = r * (1.0 - Fi), E = r * r * Fi;
            double mm = 1.0 - x; if (mm < 0) mm = 0;
            double m2 = mm * mm, m4 = m2 * m2, Fo = m4 * mm;
            double g = C + Dd * y + E * y * y;
            double K = x * (A + B * y + Fo * g);
            double wd = nd->wq;
            if (level > 0) {
                double gy = Dd + 2 * E * y, gyy = 2 * E;
                double Fo1 = -5 * m4, Fo2 = 20 * m2 * mm;
                double Kx = (A + B * y + Fo * g) + x * Fo1 * g;
                double Ky = x * (B + Fo * gy);
                double Kxx = straddle ? 0.0 : 2 * Fo1 * g + x * Fo2 * g;
                double Kxy = B + Fo * gy + x * Fo1 * gy;
                double Kyy = x * Fo * gyy;
                double vSv = v[0] * v[0] * S[0] + v[1] * v[1] * S[1] + v[2] * v[2] * S[2] +
                             2 * (v[0] * v[1] * S[3] + v[0] * v[2] * S[4] + v[1] * v[2] * S[5]);
                double nSv = n[0] * v[0] * S[0] + n[1] * v[1] * S[1] + n[2] * v[2] * S[2] +
                             (n[0] * v[1] + n[1] * v[0]) * S[3] + (n[0] * v[2] + n[2] * v[0]) * S[4] +
                             (n[1] * v[2] + n[2] * v[1]) * S[5];
                K += 0.5 * (Kxx * nSn + 2 * Kxy * nSv + Kyy * vSv);
                if (rc->chan)
                    for (int c = 0; c < 3; ++c) {
                        double nm = n[0] * R->dc[c][0] + n[1] * R->dc[c][1] + n[2] * R->dc[c][2];
                        double vm = v[0] * R->dc[c][0] + v[1] * R->dc[c][1] + v[2] * R->dc[c][2];
                        acc[c] += wd * rc->cdiff[c] * (Kx * nm + Ky * vm);
                    }
            }
            wd *= K;
            acc[0] += wd * rc->cdiff[0] * R->e[0]; acc[1] += wd * rc->cdiff[1] * R->e[1]; acc[2] += wd * rc->cdiff[2] * R->e[2];
        }
    }
}

typedef struct { int* v; int n, cap; } List;
static void push(List* l, int x) {
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 256; l->v = (int*)realloc(l->v, sizeof(int) * l->cap); }
    l->v[l->n++] = x;
}

/* one block of nodes of pass p: rows [I0,I1) x cols [J0,J1) of the lattice-2^p node grid */
static void process_block(const Pyr* P, const RC* rc, int flip, int p, int I0, int I1, int J0, int J1, const List* in,
                          double* out, Stats* st) {
    int nn = (I1 - I0) * (J1 - J0);
    Node* nodes = (Node*)malloc(sizeof(Node) * nn);
    int k = 0;
    for (int I = I0; I < I1; ++I)
        for (int J = J0; J < J1; ++J) make_node(rc, p, I, J, flip, &nodes[k++]);
    /* cone of the block's nodes */
    double ax[3] = {0, 0, 0};
    for (int i = 0; i < nn; ++i) for (int c = 0; c < 3; ++c) ax[c] += nodes[i].n[c];
    double al = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    for (int c = 0; c < 3; ++c) ax[c] /= al;
    double beta = 0;
    for (int i = 0; i < nn; ++i) {
        double dx = nodes[i].n[0] - ax[0], dy = nodes[i].n[1] - ax[1], dz = nodes[i].n[2] - ax[2];
        double ang = 2 * asin(fmin(1.0, 0.5 * sqrt(dx * dx + dy * dy + dz * dz)));
        if (ang > beta) beta = ang;
    }
    /* nodes stand for their sub-cells: widen by half a sub-cell diagonal */
    beta += 0.75 * rc->cell / (1 << p);
    double* acc = (double*)calloc((size_t)nn * 3, sizeof(double));
    List stack = {0, 0, 0}, hand = {0, 0, 0};
    for (int i = 0; i < in->n; ++i) push(&stack, in->v[i]);
    while (stack.n) {
        int ent = stack.v[--stack.n];
        int lev = ent >> 24 & 15, id = ent & 0xffffff;
        /* ids above 2^24 need a wider entry; fine for the prototype sizes (level 0 of 2000x1000 = 2e6) */
        const Rec* R = &P->lev[lev][id];
        st->visits++;
        int has_s = (rc->terms & 1) && (R->w[0] + R->w[1] + R->w[2] > 0);
        int has_d = (rc->terms & 2) && (R->e[0] + R->e[1] + R->e[2] > 0);
        if (!has_s && !has_d) continue;
        /* direction of the cell and its radius: visibility cull */
        double dn = sqrt(R->md[0] * R->md[0] + R->md[1] * R->md[1] + R->md[2] * R->md[2]);
        double dc[3] = {R->md[0] / dn, R->md[1] / dn, R->md[2] / dn};
        double rda = 2 * asin(fmin(1.0, 0.5 * (R->rd + (1 - dn))));
        double adc = ax[0] * dc[0] + ax[1] * dc[1] + ax[2] * dc[2];
        double spread = beta + rda + 1e-3;
        if (spread < 1.5607 && adc <= -sin(spread)) continue;
        int accept = 1, handover = 0;
        if (has_s) {
            double mn = sqrt(R->mu[0] * R->mu[0] + R->mu[1] * R->mu[1] + R->mu[2] * R->mu[2]);
            double hx = R->mu[0] / mn - ax[0], hy = R->mu[1] / mn - ax[1], hz = R->mu[2] / mn - ax[2];
            double ang = 2 * asin(fmin(1.0, 0.5 * sqrt(hx * hx + hy * hy + hz * hz)));
            double rha = 2 * asin(fmin(1.0, 0.5 * (R->rh + (1 - mn))));
            double dmin = ang - beta - rha; if (dmin < 0) dmin = 0;
            if (p < rc->pk && dmin < rc->thr[p]) {
                /* near for this lattice: small cells go to the child blocks, large ones are refined here so that their
                   far parts stay on this lattice */
                if (lev == 0 || rha <= rc->hand * rc->thr[p]) handover = 1; else accept = 0;
            }
            else if (lev > 0 && rha > fmin(rc->kappa * sqrt(rc->alpha2 + dmin * dmin), rc->rcap)) accept = 0;
            /* the step of the lobe at the horizon of the normals */
            if (lev > 0 && rda > rc->hz && fabs(adc) < sin(fmin(1.5607, spread))) accept = 0;
        }
        if (has_d && !handover) {
            if (lev > 0 && rda > rc->kappa_d) accept = 0;
            if (lev > 0 && rda > rc->hz && fabs(adc) < sin(fmin(1.5607, spread))) accept = 0;
        }
        if (handover) { push(&hand, ent); st->handed++; continue; }
        if (accept) {
            for (int i = 0; i < nn; ++i) eval_pair(rc, &nodes[i], R, lev, acc + 3 * i);
            if (lev == 0) st->pairs0 += nn; else st->pairs1 += nn;
        } else {
            int r = id / P->W[lev], c = id % P->W[lev];
            for (int dr = 0; dr < 2; ++dr)
                for (int dcc = 0; dcc < 2; ++dcc) {
                    int rr = 2 * r + dr, cc = 2 * c + dcc;
                    if (rr < P->H[lev - 1] && cc < P->W[lev - 1]) push(&stack, ((lev - 1) << 24) | (rr * P->W[lev - 1] + cc));
                }
        }
    }
    /* add to the cells */
    k = 0;
    for (int I = I0; I < I1; ++I)
        for (int J = J0; J < J1; ++J) {
            int i = I >> p, j = J >> p;
            double* o = out + ((size_t)i * rc->res + j) * 3;
            for (int c = 0; c < 3; ++c) {
#pragma omp atomic
                o[c] += acc[3 * k + c];
            }
            ++k;
        }
    free(nodes); free(acc); free(stack.v);
    if (hand.n) {
        /* children blocks: the four quadrants of this block's cell region, on the next lattice */
        int hI = (I1 - I0), hJ = (J1 - J0);
        for (int qa = 0; qa < 2; ++qa)
            for (int qb = 0; qb < 2; ++qb) {
                int cI0 = 2 * I0 + qa * hI, cJ0 = 2 * J0 + qb * hJ;
                process_block(P, rc, flip, p + 1, cI0, cI0 + hI, cJ0, cJ0 + hJ, &hand, out, st);
            }
    }
    free(hand.v);
}

/* gl: 5 lattices (1,2,4,8,16), each 16 doubles of nodes then 16 of weights */
int tree_render(const float* env, int He, int We, const double* z6, const double* view3, int flip, int res, int pk,
                const double* gl, double alpha_min, int terms, double kappa, double kappa_d, double hz,
                double level_scale, int pixcov, int full2, int chan, double rcap, double hand, int bh, int bw, double* out, long* stats) {
    RC rc;
    memset(&rc, 0, sizeof(rc));
    camera_frame(view3, rc.vhat, rc.left, rc.upp);
    rc.m = z6[0]; rc.rough = z6[4];
    rc.alpha = fmax(z6[4] * z6[4], alpha_min);
    rc.alpha2 = rc.alpha * rc.alpha;
    rc.inv_a2m1 = 1.0 / rc.alpha2 - 1.0;
    rc.one_m_a2 = 1.0 - rc.alpha2;
    rc.eta = 2.0 / (1.0 - sqrt(0.08 * z6[5])) - 1.0;
    for (int k = 0; k < 3; ++k) { rc.base[k] = z6[1 + k]; rc.cdiff[k] = (1.0 - rc.m) * rc.base[k] / M_PI; }
    rc.res = res; rc.cell = M_PI / res; rc.pk = pk; rc.terms = terms;
    rc.kappa = kappa; rc.kappa_d = kappa_d; rc.hz = hz; rc.pixcov = pixcov; rc.full2 = full2; rc.chan = chan; rc.rcap = rcap; rc.hand = hand;
    for (int p = 0; p < 5; ++p) { rc.glx[p] = gl + 32 * p; rc.glw[p] = gl + 32 * p + 16; }
    {
        double cell = rc.cell, alpha = rc.alpha, ca = cell * alpha;
        rc.thr[0] = 21.0 * sqrt(ca);
        rc.thr[1] = 7.5 * pow(cell, 2.0 / 3.0) * pow(alpha, 1.0 / 3.0);
        rc.thr[2] = 2.4 * pow(cell, 0.8) * pow(alpha, 0.2);
        rc.thr[3] = 1.2 * pow(cell, 8.0 / 9.0) * pow(alpha, 1.0 / 9.0);
        rc.thr[4] = 0;
        for (int i = 0; i < 4; ++i) rc.thr[i] = level_scale * fmax(rc.thr[i], 6.0 * alpha);
    }
    Pyr P;
    memset(&P, 0, sizeof(P));
    build_pyramid(&P, env, He, We, &rc);
    List top = {0, 0, 0};
    for (int i = 0; i < P.H[P.L] * P.W[P.L]; ++i) push(&top, (P.L << 24) | i);
    memset(out, 0, sizeof(double) * res * res * 3);
    int nbi = (res + bh - 1) / bh, nbj = (res + bw - 1) / bw;
    long tv = 0, t0 = 0, t1 = 0, th = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tv, t0, t1, th)
    for (int b = 0; b < nbi * nbj; ++b) {
        int bi = b / nbj, bj = b % nbj;
        Stats st = {0, 0, 0, 0};
        int I1 = (bi + 1) * bh < res ? (bi + 1) * bh : res, J1 = (bj + 1) * bw < res ? (bj + 1) * bw : res;
        process_block(&P, &rc, flip, 0, bi * bh, I1, bj * bw, J1, &top, out, &st);
        tv += st.visits; t0 += st.pairs0; t1 += st.pairs1; th += st.handed;
    }
    stats[0] = tv; stats[1] = t0; stats[2] = t1; stats[3] = th;
    for (int l = 0; l <= P.L; ++l) free(P.lev[l]);
    free(top.v);
    return 0;
}
