// MFE fold kernel (Zuker recursions, Turner-2004 nearest-neighbour model, dangles=2) for sm_100a.
//
// Replaces RNA.fold_compound(seq, md).mfe() -- ScanFold.py:494-497/:513/:541 (native window) and
// ScanFoldFunctions.py:774-789 (rna_folder, the r background folds per window).
//
// One CTA folds one sequence at a time (persistent grid, CTAs stride over the batch).  The C and FML
// matrices are stored DIAGONAL-MAJOR so that the cells of one anti-diagonal d = j - i -- which are
// independent -- sit at consecutive shared-memory addresses: a wavefront step reads diagonals < d and
// writes diagonal d.  Interior-loop candidates are split over G lanes per pairable cell and reduced
// with warp shuffles; the three separable loop classes (generic, 1xn, bulge) read rolling 32-diagonal
// buffers that already include the inner pair's mismatch term, so a candidate costs one shared-memory
// load plus one add-min.  Energies are int32 dcal, INF = 10^7, exactly as ViennaRNA.
#include <cstdio>

#include "device_common.cuh"

namespace sfb {
namespace {

constexpr int NT = 256;        // threads per CTA
constexpr int NCAND = 496;     // (u1,u2), u1+u2 <= 30
constexpr int ROLL = 33;       // rolling diagonals: reads reach back to d - 32 while diagonal d is written

enum { CLS_GENERIC = 0, CLS_1N = 1, CLS_BULGE = 2, CLS_TABLE = 3 };

struct SmallTab {
    int stack[64];
    int mmI[200], mm1n[200], mm23[200], mmM[200], mmExt[200], mmH[200];
    int d5[40], d3[40];
    int bulge[31], il[31];
    int MLbase, MLclosing, MLintern, ninio, max_ninio, TerminalAU;
    int cand[NCAND];      // u1 | u2<<5 | cls<<10 | size<<16, sorted by u1+u2
    int ncand_upto[32];   // number of candidates with u1+u2 <= u
};

struct FoldCtx {
    const MfeTables *T;   // global tables (int11/int21/int22, hairpin_len, special loops)
    const SmallTab *st;   // shared
    const uint8_t *S;     // shared, codes, W entries
    const int32_t *sc;    // shared 1-based or NULL
    int W;
    HcCtx h;              // pair permission (hard constraints)
};

__device__ __forceinline__ int mm_idx(int t, int a, int b) { return (t * 5 + a) * 5 + b; }

__device__ int e_hairpin(const FoldCtx &c, int i, int j, int type) {
    int u = j - i - 1;
    int e = c.T->hairpin_len[u];
    if (u < 3) return e;
    if (u == 4) {
        int key = loop_key_dev(c.S, i, 6);
        for (int k = 0; k < c.T->n_tetra; k++)
            if (c.T->tetra_key[k] == key) return c.T->tetra_e[k];
    } else if (u == 6) {
        int key = loop_key_dev(c.S, i, 8);
        for (int k = 0; k < c.T->n_hexa; k++)
            if (c.T->hexa_key[k] == key) return c.T->hexa_e[k];
    } else if (u == 3) {
        int key = loop_key_dev(c.S, i, 5);
        for (int k = 0; k < c.T->n_tri; k++)
            if (c.T->tri_key[k] == key) return c.T->tri_e[k];
        return e + (type > 2 ? c.st->TerminalAU : 0);
    }
    return e + c.st->mmH[mm_idx(type, c.S[i + 1], c.S[j - 1])];
}

// full interior-loop energy, all classes (used for the table classes in the fill and for traceback)
__device__ int e_intloop(const FoldCtx &c, int n1, int n2, int type, int t2, int si1, int sj1, int sp1, int sq1) {
    const SmallTab &s = *c.st;
    int nl = max(n1, n2), ns = min(n1, n2);
    if (nl == 0) return s.stack[type * 8 + t2];
    if (ns == 0) {
        int e = s.bulge[nl];
        if (nl == 1)
            e += s.stack[type * 8 + t2];
        else {
            if (type > 2) e += s.TerminalAU;
            if (t2 > 2) e += s.TerminalAU;
        }
        return e;
    }
    if (ns == 1) {
        if (nl == 1) return c.T->int11[type][t2][si1][sj1];
        if (nl == 2) {
            if (n1 == 1) return c.T->int21[type][t2][si1][sq1][sj1];
            return c.T->int21[t2][type][sq1][si1][sp1];
        }
        return s.il[nl + 1] + min(s.max_ninio, (nl - ns) * s.ninio) + s.mm1n[mm_idx(type, si1, sj1)] +
               s.mm1n[mm_idx(t2, sq1, sp1)];
    }
    if (ns == 2) {
        if (nl == 2) return c.T->int22[type][t2][si1][sp1][sq1][sj1];
        if (nl == 3) return s.il[5] + s.ninio + s.mm23[mm_idx(type, si1, sj1)] + s.mm23[mm_idx(t2, sq1, sp1)];
    }
    return s.il[nl + ns] + min(s.max_ninio, (nl - ns) * s.ninio) + s.mmI[mm_idx(type, si1, sj1)] +
           s.mmI[mm_idx(t2, sq1, sp1)];
}

__device__ __forceinline__ int e_mlstem(const SmallTab &s, int type, int si1, int sj1) {
    int e = 0;
    if (si1 >= 0 && sj1 >= 0)
        e = s.mmM[mm_idx(type, si1, sj1)];
    else if (si1 >= 0)
        e = s.d5[type * 5 + si1];
    else if (sj1 >= 0)
        e = s.d3[type * 5 + sj1];
    if (type > 2) e += s.TerminalAU;
    return e + s.MLintern;
}

__device__ __forceinline__ int e_extloop(const SmallTab &s, int type, int si1, int sj1) {
    int e = 0;
    if (si1 >= 0 && sj1 >= 0)
        e = s.mmExt[mm_idx(type, si1, sj1)];
    else if (si1 >= 0)
        e = s.d5[type * 5 + si1];
    else if (sj1 >= 0)
        e = s.d3[type * 5 + sj1];
    if (type > 2) e += s.TerminalAU;
    return e;
}

__device__ __forceinline__ int sc_stack4(const FoldCtx &c, int i, int j, int p, int q) {
    return c.sc ? c.sc[i + 1] + c.sc[p + 1] + c.sc[q + 1] + c.sc[j + 1] : 0;
}

// serial traceback (one thread), candidate order of SURVEY A.4; pair table is 1-based partner, 0 = unpaired
__device__ bool traceback(const FoldCtx &c, const int *C, const int *M, const int *F, int *stk, int16_t *pt) {
    const int W = c.W;
    const SmallTab &s = *c.st;
    for (int k = 0; k < W; k++) pt[k] = 0;
    int sp = 0;
    stk[0] = 0;
    stk[1] = W - 1;
    stk[2] = 0;
    sp = 1;
#define CC(i, j) C[tri_off((j) - (i), W) + (i)]
#define MM(i, j) (((j) - (i)) > TURN ? M[tri_off((j) - (i), W) + (i)] : INF)
    while (sp > 0) {
        sp--;
        int i = stk[3 * sp], j = stk[3 * sp + 1], ml = stk[3 * sp + 2];
        bool have_pair = false;
        if (j < i + TURN + 1) continue;
        int fij = ml ? MM(i, j) : F[j + 1];
        int mij1 = MM(i, j - 1);
        int fi = ml ? (mij1 < INF ? mij1 + s.MLbase : INF) : F[j];
        if (fij == fi) {
            stk[3 * sp] = i;
            stk[3 * sp + 1] = j - 1;
            stk[3 * sp + 2] = ml;
            sp++;
            continue;
        }
        if (ml == 0) {
            int k;
            bool found = false;
            for (k = j - TURN - 1; k >= 0; k--) {
                int ckj = CC(k, j);
                if (ckj >= INF) continue;
                int type = pair_type(c.S[k], c.S[j]);
                if (fij == e_extloop(s, type, k > 0 ? c.S[k - 1] : -1, j < W - 1 ? c.S[j + 1] : -1) + ckj + F[k]) {
                    found = true;
                    break;
                }
            }
            if (!found) return false;
            stk[3 * sp] = 0;
            stk[3 * sp + 1] = k - 1;
            stk[3 * sp + 2] = 0;
            sp++;
            i = k;
            have_pair = true;
        } else {
            int mi1j = MM(i + 1, j);
            if (mi1j < INF && mi1j + s.MLbase == fij) {
                stk[3 * sp] = i + 1;
                stk[3 * sp + 1] = j;
                stk[3 * sp + 2] = 1;
                sp++;
                continue;
            }
            int cij = CC(i, j);
            if (cij < INF &&
                fij == cij + e_mlstem(s, pair_type(c.S[i], c.S[j]), i > 0 ? c.S[i - 1] : -1, j < W - 1 ? c.S[j + 1] : -1)) {
                have_pair = true;
            } else {
                int k;
                for (k = i + 1 + TURN; k <= j - 2 - TURN; k++) {
                    int a = MM(i, k), b = MM(k + 1, j);
                    if (a < INF && b < INF && fij == a + b) break;
                }
                if (k > j - 2 - TURN) return false;
                stk[3 * sp] = i;
                stk[3 * sp + 1] = k;
                stk[3 * sp + 2] = 1;
                sp++;
                stk[3 * sp] = k + 1;
                stk[3 * sp + 1] = j;
                stk[3 * sp + 2] = 1;
                sp++;
                continue;
            }
        }
        while (have_pair) {
            pt[i] = (int16_t)(j + 1);
            pt[j] = (int16_t)(i + 1);
            int type = pair_type(c.S[i], c.S[j]);
            int cij = CC(i, j);
            if (cij == e_hairpin(c, i, j, type)) break;
            bool traced = false;
            int pmax = min(j - 2 - TURN, i + MAXLOOP + 1);
            for (int p = i + 1; p <= pmax && !traced; p++) {
                int minq = j - i + p - MAXLOOP - 2;
                if (minq < p + 1 + TURN) minq = p + 1 + TURN;
                for (int q = j - 1; q >= minq; q--) {
                    int cpq = CC(p, q);
                    if (cpq >= INF) continue;
                    int t2 = rtype_of(pair_type(c.S[p], c.S[q]));
                    int e = e_intloop(c, p - i - 1, j - q - 1, type, t2, c.S[i + 1], c.S[j - 1], c.S[p - 1], c.S[q + 1]);
                    if (p == i + 1 && q == j - 1) e += sc_stack4(c, i, j, p, q);
                    if (cij == e + cpq) {
                        i = p;
                        j = q;
                        traced = true;
                        break;
                    }
                }
            }
            if (traced) continue;
            int en = cij - e_mlstem(s, rtype_of(type), c.S[j - 1], c.S[i + 1]) - s.MLclosing;
            int k;
            for (k = i + 2 + TURN; k < j - 2 - TURN; k++) {
                int a = MM(i + 1, k), b = MM(k + 1, j - 1);
                if (a < INF && b < INF && en == a + b) break;
            }
            if (k > j - 3 - TURN) return false;
            stk[3 * sp] = i + 1;
            stk[3 * sp + 1] = k;
            stk[3 * sp + 2] = 1;
            sp++;
            stk[3 * sp] = k + 1;
            stk[3 * sp + 1] = j - 1;
            stk[3 * sp + 2] = 1;
            sp++;
            break;
        }
    }
#undef CC
#undef MM
    return true;
}

__device__ __forceinline__ int floor_pow2(int x) { return x <= 1 ? 1 : 1 << (31 - __clz(x)); }

#ifndef SFB_MFE1_MINB
#define SFB_MFE1_MINB 6   // resident CTAs per SM (global-memory mode, W=600): 3 -> 3.1 k, 4 -> 3.9 k, 6 -> 4.5 k folds/s
#endif
__global__ void __launch_bounds__(NT, SFB_MFE1_MINB) mfe_fold_kernel(MfeLaunch L, const MfeTables *__restrict__ T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = L.W;
    const int tid = threadIdx.x, lane = tid & 31;
    const int ntri = W * (W + 1) / 2;

    // ---- shared memory carve-up
    SmallTab *st = reinterpret_cast<SmallTab *>(smem_raw);
    int *ip = reinterpret_cast<int *>(smem_raw + ((sizeof(SmallTab) + 15) & ~15));
    int *F = ip;            ip += W + 1;
    int *DML = ip;          ip += 3 * W;
    int *scs = ip;          ip += W + 1;
    int *tbstk = ip;        ip += 3 * (2 * W + 8);
    int *misc = ip;         ip += 4;   // [0] list count
    int *roll, *Cm, *Mm;
    if (L.mats_in_gmem == 0) {
        roll = ip;          ip += 3 * ROLL * W;
        Cm = ip;            ip += ntri;
        Mm = ip;            ip += ntri;
    } else {
        int *g = L.gscratch + (long long)blockIdx.x * L.gscratch_per_cta;
        if (L.mats_in_gmem == 1) {
            roll = ip;      ip += 3 * ROLL * W;
        } else {
            roll = g;       g += 3 * ROLL * W;
        }
        Cm = g;             g += ntri;
        Mm = g;
    }
    int16_t *list = reinterpret_cast<int16_t *>(ip);
    int16_t *mate = list + W;
    int16_t *pt = mate + W;
    uint8_t *S = reinterpret_cast<uint8_t *>(pt + W);
    uint8_t *hcf = S + W + 4;
    uint8_t *ctype = hcf + W + 4;

    // ---- stage the small tables once per CTA
    {
        auto cp = [&](int *dst, const int *s, int n) {
            for (int k = tid; k < n; k += NT) dst[k] = s[k];
        };
        cp(st->stack, &T->stack[0][0], 64);
        cp(st->mmI, &T->mismatchI[0][0][0], 200);
        cp(st->mm1n, &T->mismatch1nI[0][0][0], 200);
        cp(st->mm23, &T->mismatch23I[0][0][0], 200);
        cp(st->mmM, &T->mismatchM[0][0][0], 200);
        cp(st->mmExt, &T->mismatchExt[0][0][0], 200);
        cp(st->mmH, &T->mismatchH[0][0][0], 200);
        cp(st->d5, &T->dangle5[0][0], 40);
        cp(st->d3, &T->dangle3[0][0], 40);
        cp(st->bulge, T->bulge, 31);
        cp(st->il, T->internal_loop, 31);
        if (tid == 0) {
            st->MLbase = T->MLbase;
            st->MLclosing = T->MLclosing;
            st->MLintern = T->MLintern;
            st->ninio = T->ninio;
            st->max_ninio = T->max_ninio;
            st->TerminalAU = T->TerminalAU;
            // candidate table sorted by total loop size u = u1 + u2
            int n = 0;
            for (int u = 0; u <= MAXLOOP; u++) {
                for (int u1 = 0; u1 <= u; u1++) {
                    int u2 = u - u1, cls, size = 0;
                    int nl = max(u1, u2), ns = min(u1, u2);
                    if (ns == 0 && nl >= 2) {
                        cls = CLS_BULGE;
                        size = T->bulge[nl];
                    } else if (ns == 1 && nl >= 3) {
                        cls = CLS_1N;
                        size = T->internal_loop[nl + 1] + min(T->max_ninio, (nl - ns) * T->ninio);
                    } else if (ns >= 2 && !(ns == 2 && nl <= 3)) {
                        cls = CLS_GENERIC;
                        size = T->internal_loop[u] + min(T->max_ninio, (nl - ns) * T->ninio);
                    } else {
                        cls = CLS_TABLE;
                    }
                    st->cand[n++] = u1 | (u2 << 5) | (cls << 10) | (size << 16);
                }
                st->ncand_upto[u] = n;
            }
        }
    }
    __syncthreads();

    FoldCtx c;
    c.T = T;
    c.st = st;
    c.S = S;
    c.W = W;
    c.h.S = S;
    c.h.W = W;
    c.h.max_span = L.max_span;

    for (int fold = blockIdx.x; fold < L.n_fold; fold += gridDim.x) {
        if (L.redo_only && L.e_out[fold] != MFE_REDO) continue;
        // ---- per-fold prologue
        for (int k = tid; k < W; k += NT) S[k] = L.seqs[(long long)fold * W + k];
        c.h.hcf = nullptr;
        c.h.mate = nullptr;
        c.sc = nullptr;
        c.h.n_enf = 0;
        if (L.hc) {
            for (int k = tid; k < W; k += NT) {
                char ch = (char)L.hc[(long long)fold * W + k];
                hcf[k] = (ch == 'x' ? 1 : 0) | (ch == '<' ? 2 : 0) | (ch == '>' ? 4 : 0);
                mate[k] = -1;
            }
            c.h.hcf = hcf;
            c.h.mate = mate;
        }
        if (L.sc) {
            for (int k = tid; k <= W; k += NT) scs[k] = L.sc[(long long)fold * (W + 1) + k];
            c.sc = scs;
        }
        for (int k = tid; k < 3 * ROLL * W; k += NT) roll[k] = INF;
        for (int k = tid; k < 3 * W; k += NT) DML[k] = INF;
        __syncthreads();
        if (L.hc && tid == 0) {  // match brackets (weak enforcement); unbalanced ones are ignored
            int sp = 0, n_enf = 0;
            for (int k = 0; k < W; k++) {
                char ch = (char)L.hc[(long long)fold * W + k];
                if (ch == '(')
                    tbstk[sp++] = k;
                else if (ch == ')' && sp > 0) {
                    int a = tbstk[--sp];
                    mate[a] = (int16_t)k;
                    mate[k] = (int16_t)a;
                    n_enf++;
                }
            }
            misc[1] = n_enf;
        }
        __syncthreads();
        if (L.hc) c.h.n_enf = misc[1];

        // ---- wavefront over anti-diagonals
        for (int d = TURN + 1; d < W; d++) {
            const int ncells = W - d;
            const int tri_d = tri_off(d, W);
            const int rrow = (d % ROLL) * W;
            if (tid == 0) misc[0] = 0;
            __syncthreads();
            // phase 1a: pair permission per cell, compact the pairable ones
            for (int i = tid; i < ncells; i += NT) {
                int t = allowed_type(c.h, i, i + d);
                ctype[i] = (uint8_t)t;
                if (!t) {
                    Cm[tri_d + i] = INF;
                    roll[rrow + i] = INF;
                    roll[ROLL * W + rrow + i] = INF;
                    roll[2 * ROLL * W + rrow + i] = INF;
                } else {
                    int pos = atomicAdd(&misc[0], 1);
                    list[pos] = (int16_t)i;
                }
            }
            __syncthreads();
            // phase 1b: C[i,j] for pairable cells; G lanes share one cell's interior-loop candidates
            {
                const int ncp = misc[0];
                const int G = min(32, floor_pow2(NT / max(ncp, 1)));
                const int gsh = 31 - __clz(G);
                const int items = ncp << gsh;
                const int umax = min(MAXLOOP, d - 2 - (TURN + 1));
                const int ncand = umax >= 0 ? st->ncand_upto[umax] : 0;
                for (int base = tid - lane; base < items; base += NT) {
                    const int item = base + lane;
                    const bool active = item < items;
                    int acc = INF, i = 0, j = 0, type = 0;
                    if (active) {
                        i = list[item >> gsh];
                        j = i + d;
                        type = ctype[i];
                        const int g = item & (G - 1);
                        const int si1 = S[i + 1], sj1 = S[j - 1];
                        const int mi = mm_idx(type, si1, sj1);
                        const int outer0 = st->mmI[mi], outer1 = st->mm1n[mi];
                        const int outer2 = type > 2 ? st->TerminalAU : 0;
                        for (int ci = g; ci < ncand; ci += G) {
                            const int cd = st->cand[ci];
                            const int u1 = cd & 31, u2 = (cd >> 5) & 31, cls = (cd >> 10) & 7;
                            const int p = i + 1 + u1, dd = d - 2 - u1 - u2;
                            int v;
                            if (cls != CLS_TABLE) {
                                const int outer = cls == CLS_GENERIC ? outer0 : (cls == CLS_1N ? outer1 : outer2);
                                v = roll[cls * ROLL * W + (dd % ROLL) * W + p] + (cd >> 16) + outer;
                            } else {
                                const int q = j - 1 - u2;
                                const int cpq = Cm[tri_off(dd, W) + p];
                                v = INF;
                                if (cpq < INF) {
                                    const int t2 = rtype_of(pair_type(S[p], S[q]));
                                    v = cpq + e_intloop(c, u1, u2, type, t2, si1, sj1, S[p - 1], S[q + 1]);
                                    if (u1 == 0 && u2 == 0) v += sc_stack4(c, i, j, p, q);
                                }
                            }
                            acc = min(acc, v);
                        }
                    }
                    for (int o = 1; o < G; o <<= 1) acc = min(acc, __shfl_xor_sync(0xffffffffu, acc, o));
                    if (active && (item & (G - 1)) == 0) {
                        int e = min(acc, e_hairpin(c, i, j, type));
                        if (d >= 2 + TURN + 1 + 2) {  // multiloop closed by (i,j): split-min of cell (i+1,j-1)
                            const int dm = DML[((d - 2) % 3) * W + i + 1];
                            if (dm < INF)
                                e = min(e, dm + e_mlstem(*st, rtype_of(type), S[j - 1], S[i + 1]) + st->MLclosing);
                        }
                        Cm[tri_d + i] = e;
                        int vg = INF, v1 = INF, vb = INF;
                        if (i > 0 && j < W - 1) {  // (i,j) as the inner pair of an enclosing loop
                            const int t2 = rtype_of(type);
                            const int m2 = mm_idx(t2, S[j + 1], S[i - 1]);
                            vg = e + st->mmI[m2];
                            v1 = e + st->mm1n[m2];
                            vb = e + (t2 > 2 ? st->TerminalAU : 0);
                        }
                        roll[rrow + i] = vg;
                        roll[ROLL * W + rrow + i] = v1;
                        roll[2 * ROLL * W + rrow + i] = vb;
                    }
                }
            }
            __syncthreads();
            // phase 2: FML[i,j] for every cell; G2 lanes share the split loop
            {
                const int G = min(32, floor_pow2(NT / ncells));
                const int gsh = 31 - __clz(G);
                const int items = ncells << gsh;
                const int tri_d1 = tri_off(d - 1, W);
                for (int base = tid - lane; base < items; base += NT) {
                    const int item = base + lane;
                    const bool active = item < items;
                    int dec = INF, i = 0;
                    if (active) {
                        i = item >> gsh;
                        for (int k = TURN + 1 + (item & (G - 1)); k <= d - 2 - TURN; k += G) {
                            const int a = Mm[tri_off(k, W) + i];
                            const int b = Mm[tri_off(d - k - 1, W) + i + k + 1];
                            dec = min(dec, a + b);
                        }
                    }
                    for (int o = 1; o < G; o <<= 1) dec = min(dec, __shfl_xor_sync(0xffffffffu, dec, o));
                    if (active && (item & (G - 1)) == 0) {
                        const int j = i + d;
                        if (dec > INF / 2) dec = INF;
                        int m = dec;
                        if (d - 1 > TURN) {
                            const int a = Mm[tri_d1 + i + 1], b = Mm[tri_d1 + i];
                            if (a < INF) m = min(m, a + st->MLbase);
                            if (b < INF) m = min(m, b + st->MLbase);
                        }
                        const int cij = Cm[tri_d + i];
                        if (cij < INF)
                            m = min(m, cij + e_mlstem(*st, ctype[i], i > 0 ? S[i - 1] : -1, j < W - 1 ? S[j + 1] : -1));
                        DML[(d % 3) * W + i] = dec;
                        Mm[tri_d + i] = m;
                    }
                }
            }
            __syncthreads();
        }

        // ---- exterior loop F[len], one warp, sequential in len, lanes over the 5' end
        if (tid < 32) {
            for (int k = lane; k <= min(W, TURN + 1); k += 32) F[k] = 0;
            __syncwarp();
            for (int len = TURN + 2; len <= W; len++) {
                const int j = len - 1;
                int best = INF;
                for (int i = lane; i <= j - TURN - 1; i += 32) {
                    const int cij = Cm[tri_off(j - i, W) + i];
                    if (cij < INF) {
                        const int t = pair_type(S[i], S[j]);
                        best = min(best, F[i] + cij + e_extloop(*st, t, i > 0 ? S[i - 1] : -1, j < W - 1 ? S[j + 1] : -1));
                    }
                }
                for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                if (lane == 0) F[len] = min(F[len - 1], best);
                __syncwarp();
            }
            if (lane == 0) L.e_out[fold] = F[W];
        }
        __syncthreads();
        if (L.pair_tbl) {
            if (tid == 0) {
                bool ok = traceback(c, Cm, Mm, F, tbstk, pt);
                if (!ok) L.e_out[fold] = INF;  // surfaces as an error on the host
            }
            __syncthreads();
            for (int k = tid; k < W; k += NT) L.pair_tbl[(long long)fold * W + k] = pt[k];
        }
        __syncthreads();
    }
}

size_t mfe_smem_bytes(int W, int mats_in_gmem) {
    size_t ints = (W + 1) + 3 * W + (W + 1) + 3 * (2 * W + 8) + 4;
    const size_t ntri = (size_t)W * (W + 1) / 2;
    if (mats_in_gmem == 0) ints += 3 * ROLL * W + 2 * ntri;
    if (mats_in_gmem == 1) ints += 3 * ROLL * W;
    size_t bytes = ((sizeof(SmallTab) + 15) & ~15) + ints * 4;
    bytes += 3 * W * sizeof(int16_t) + 3 * (W + 4);
    return (bytes + 15) & ~(size_t)15;
}

}  // namespace

size_t mfe_scratch_ints_per_cta(int W, int *mats_in_gmem) {
    const size_t limit = 227 * 1024;
    const size_t ntri = (size_t)W * (W + 1) / 2;
    int mode = 0;
    if (mfe_smem_bytes(W, 0) > limit) mode = 1;
    if (mode == 1 && mfe_smem_bytes(W, 1) > limit) mode = 2;
#ifndef SFB_MFE1_GMEM_OCC
#define SFB_MFE1_GMEM_OCC 3   // measured: W=260 22.5 k -> 48.6 k, W=400 6.4 k -> 15.9 k folds/s with the buffers in global memory
#endif
    {
    // rolling buffers in shared memory leave room for few CTAs per SM; in global memory the kernel keeps SFB_MFE1_MINB
    if (mode == 1 && (227 * 1024) / (mfe_smem_bytes(W, 1) + 1024) < SFB_MFE1_GMEM_OCC) mode = 2;
    }
    if (mats_in_gmem) *mats_in_gmem = mode;
    if (mode == 0) return 0;
    return 2 * ntri + (mode == 2 ? 3 * ROLL * (size_t)W : 0);
}

int mfe_grid_size(int W, int n_sm, int n_fold) {
    int mode;
    mfe_scratch_ints_per_cta(W, &mode);
    size_t smem = mfe_smem_bytes(W, mode);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > SFB_MFE1_MINB) per_sm = SFB_MFE1_MINB;   // registers: resident CTAs per SM
    long long g = (long long)n_sm * per_sm;
    if (g > n_fold) g = n_fold;
    return (int)g;
}

void launch_mfe(const MfeLaunch &L, const MfeTables *d_tab, int n_sm, cudaStream_t stream, int *n_launches) {
    if (L.n_fold <= 0) return;
    size_t smem = mfe_smem_bytes(L.W, L.mats_in_gmem);
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(mfe_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        configured = 227 * 1024;
    }
    int grid = mfe_grid_size(L.W, n_sm, L.n_fold);
    mfe_fold_kernel<<<grid, NT, smem, stream>>>(L, d_tab);
    if (n_launches) (*n_launches)++;
}

}  // namespace sfb
