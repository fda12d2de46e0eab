// Partition function, second generation: one CTA (512 threads) per window, everything the inner loops touch in
// shared memory, windows up to 120 nt without constraints.
//
// Replaces fc.pf(), fc.centroid(), fc.mean_bp_distance() -- ScanFold.py:498,503-504 -- for the unconstrained
// native windows of a scan, also under per-nucleotide hard constraints ('x' '<' '>'); pf.cu keeps enforced pairs,
// soft constraints and the long windows.
//
// Both passes walk the matrix by COLUMN (3' end) instead of by anti-diagonal: every cell of a column only depends
// on earlier columns, all lanes of a warp share the column, and an interior-loop candidate (u1, u2) is then a
// load at  ring[column -+ u2][lane +- u1]  -- a compile-time row and offset -- followed by one DFMA whose size
// factor is a constant-bank operand.  The 32-column rings hold qb (inside) / the outside weight P (outside)
// already multiplied by the mismatch factor of that pair for the three separable loop classes; rows are zero
// padded so no candidate needs a bounds test.  qm lives in a folded triangular matrix with an odd pitch (row and
// column walks are both bank-conflict free); the multiloop sums with a geometric weight are prefix / suffix scans.
// Only qb (read back once per cell) and the multiloop closing weights PM (read as rows) stream through L2.
#include <cstdlib>
#include <type_traits>

#include "device_common.cuh"

namespace sfb {
namespace {

constexpr int P2 = 120;        // longest window
constexpr int RP = P2 + 32;    // ring pitch (zero padded)
constexpr int PQ = 121;        // pitch of the folded qm matrix (odd)
constexpr int QROWS = (P2 + 3) / 2 + 1;
constexpr int NT2 = 512;

__constant__ double c2_G[32 * 32];  // [u1 * 32 + u2]: expinternal[u] * expninio[|u1-u2|] * scale[u+2]
__constant__ double c2_1[32];       // 1xn loops of total size u
__constant__ double c2_B[32];       // bulges of size u

// the small Boltzmann-factor tables the narrow column phases look up (same member names as PfTables): a shared-memory
// copy per CTA, so those phases wait for an LDS instead of an L1-missing global load
struct PfHead {
    double expmismatchI[8][5][5], expmismatch1nI[8][5][5], expmismatchM[8][5][5], expmismatchExt[8][5][5];
    double expdangle5[8][5], expdangle3[8][5];
    double expMLintern, expTermAU;
};

struct Smem2 {
    PfHead th;
    double ring[3][32][RP];     // generic | 1xn | bulge copies
    double ringq[8][P2];        // raw qb (inside) / raw P (outside) of the last columns: table-driven shapes
    double qm[QROWS * PQ];      // folded: row = 3' end k', entries i <= k'-4
    double partA[4][P2 + 8];    // interior-loop partial sums by u2 group
    double partC[4][P2 + 8];    // multiloop partial sums
    double qm1[2][P2 + 8];
    double qqcol[P2 + 8];       // sum_k qm[i,k-1] qm1[k,j] of the previous column (multiloop closing)
    double ecol[P2 + 8];
    double x1[P2 + 8], x2[P2 + 8], x12[P2 + 8], g1[P2 + 8];
    double q5[P2 + 8], q3[P2 + 8];
    double qcol[P2 + 8];        // outside: qb of the current column (read from L2 once, in the wide phase)
    double pmcol[P2 + 8];       // outside: PM of the previous column
    double scale[P2 + 40], emlb[P2 + 8], ainv[P2 + 8];
    double red[32];
    short cen[P2 + 8];
    unsigned char S[P2 + 8];
    unsigned char ty[P2 + 8];   // pair type of the cells of the current column
    unsigned char can5[P2 + 8], can3[P2 + 8];   // hard constraints: may be the 5' / 3' partner of a pair
};

template <int A, int B, class F>
__device__ __forceinline__ void sfor2(F &&f) {
    if constexpr (A <= B) {
        f(std::integral_constant<int, A>{});
        sfor2<A + 1, B>(f);
    }
}

// separable interior-loop candidates with u2 = S4, S4+4, ...: inside reads ring[col0 - u2][pos + u1], outside
// ring[col0 + u2][pos - u1]
template <int S4, bool OUT>
__device__ __forceinline__ void cand_group(const double *ring, int col0, int pos, int u2max, double &aG, double &a1,
                                           double &aB) {
    double aG1 = 0.;
    sfor2<0, 7>([&](auto Q) {
        constexpr int u2 = S4 + 4 * decltype(Q)::value;
        if constexpr (u2 <= MAXLOOP) {
            if (u2 <= u2max) {
                const int col = OUT ? col0 + u2 : col0 - u2;
                const double *r = ring + (col & 31) * RP + pos;
                sfor2<0, MAXLOOP - u2>([&](auto V) {
                    constexpr int u1 = decltype(V)::value;
                    constexpr int us = u1 < u2 ? u1 : u2, ul = u1 < u2 ? u2 : u1;
                    constexpr int off = OUT ? -u1 : u1;
                    if constexpr (us == 0 && ul >= 2)
                        aB = fma(r[2 * 32 * RP + off], c2_B[ul], aB);
                    else if constexpr (us == 1 && ul >= 3)
                        a1 = fma(r[32 * RP + off], c2_1[u1 + u2], a1);
                    else if constexpr (us >= 2 && !(us == 2 && ul <= 3)) {
                        if constexpr (u1 & 1)
                            aG1 = fma(r[off], c2_G[u1 * 32 + u2], aG1);
                        else
                            aG = fma(r[off], c2_G[u1 * 32 + u2], aG);
                    }
                });
            }
        }
    });
    aG += aG1;
}

struct Ctx2 {
    const PfTables *T;
    const MfeTables *M;
    const unsigned char *S;
    const double *scale;
    int W;
};

__device__ double hairpin2(const Ctx2 &c, int i, int j, int type) {
    const int u = j - i - 1;
    const double z = c.T->exphairpin_len[u];
    if (u < 3) return z;
    if (u == 4) {
        const int key = loop_key_dev(c.S, i, 6);
        for (int k = 0; k < c.M->n_tetra; k++)
            if (c.M->tetra_key[k] == key) return c.T->exptetra[k];
    } else if (u == 6) {
        const int key = loop_key_dev(c.S, i, 8);
        for (int k = 0; k < c.M->n_hexa; k++)
            if (c.M->hexa_key[k] == key) return c.T->exphexa[k];
    } else if (u == 3) {
        const int key = loop_key_dev(c.S, i, 5);
        for (int k = 0; k < c.M->n_tri; k++)
            if (c.M->tri_key[k] == key) return c.T->exptri[k];
        return type > 2 ? z * c.T->expTermAU : z;
    }
    return z * c.T->expmismatchH[type][c.S[i + 1]][c.S[j - 1]];
}

// the nine table-driven shapes, Boltzmann factor without the length scaling
__device__ __forceinline__ double shape2(const PfTables *T, int u1, int u2, int type, int t2, int si1, int sj1, int sp1,
                                         int sq1) {
    const int ul = max(u1, u2), us = min(u1, u2);
    if (ul == 0) return T->expstack[type][t2];
    if (us == 0) return T->expbulge[1] * T->expstack[type][t2];   // bulge of one
    if (us == 1) {
        if (ul == 1) return T->expint11[type][t2][si1][sj1];
        if (u1 == 1) return T->expint21[type][t2][si1][sq1][sj1];
        return T->expint21[t2][type][sq1][si1][sp1];
    }
    if (ul == 2) return T->expint22[type][t2][si1][sp1][sq1][sj1];
    return T->expinternal[5] * T->expmismatch23I[type][si1][sj1] * T->expmismatch23I[t2][sq1][sp1] * T->expninio[1];
}

template <class TT>
__device__ __forceinline__ double mlstem2(const TT *T, int type, int si1, int sj1) {
    double z = 1.;
    if (si1 >= 0 && sj1 >= 0)
        z = T->expmismatchM[type][si1][sj1];
    else if (si1 >= 0)
        z = T->expdangle5[type][si1];
    else if (sj1 >= 0)
        z = T->expdangle3[type][sj1];
    if (type > 2) z *= T->expTermAU;
    return z * T->expMLintern;
}

template <class TT>
__device__ __forceinline__ double extloop2(const TT *T, int type, int si1, int sj1) {
    double z = 1.;
    if (si1 >= 0 && sj1 >= 0)
        z = T->expmismatchExt[type][si1][sj1];
    else if (si1 >= 0)
        z = T->expdangle5[type][si1];
    else if (sj1 >= 0)
        z = T->expdangle3[type][sj1];
    if (type > 2) z *= T->expTermAU;
    return z;
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// the nine shapes as (u1, u2) nibbles: (0,0) (0,1) (1,0) (1,1) (1,2) (2,1) (2,2) (2,3) (3,2)
__host__ __device__ constexpr int shape_u1(int z) { return (int)((0x322211100ull >> (4 * z)) & 15); }
__host__ __device__ constexpr int shape_u2(int z) { return (int)((0x232121010ull >> (4 * z)) & 15); }

// Table-driven shapes with u2 = S4 (u2 <= 3, so u2 mod 4 = u2) closed by (i,j): they ride with the separable
// candidates of the same u2 group in phase A, where all 16 warps work, instead of the 5-warp column phase.
template <int S4>
__device__ __forceinline__ double shapes_inside(const Smem2 &sm, const PfTables *T, int i, int j, int t) {
    const unsigned char *S = sm.S;
    const int si1 = S[i + 1], sj1 = S[j - 1];
    double acc = 0.;
    sfor2<0, 8>([&](auto Z) {
        constexpr int z = decltype(Z)::value, u1 = shape_u1(z), u2 = shape_u2(z);
        if constexpr (u2 == S4) {
            const int p = i + 1 + u1, q = j - 1 - u2;
            if (q - p > TURN) {
                const double qpq = sm.ringq[q & 7][p];
                if (qpq != 0.) {
                    const int t2 = rtype_of(pair_type(S[p], S[q]));
                    acc += qpq * shape2(T, u1, u2, t, t2, si1, sj1, S[p - 1], S[q + 1]) * sm.scale[u1 + u2 + 2];
                }
            }
        }
    });
    return acc;
}

// the same for the outside pass: (k,l) is the inner pair, (i,j) = (k-1-u1, l+1+u2) the closing one
template <int S4>
__device__ __forceinline__ double shapes_outside(const Smem2 &sm, const PfTables *T, int k, int l, int t2, int W) {
    const unsigned char *S = sm.S;
    const int sp1 = S[k - 1], sq1 = S[l + 1];
    double acc = 0.;
    sfor2<0, 8>([&](auto Z) {
        constexpr int z = decltype(Z)::value, u1 = shape_u1(z), u2 = shape_u2(z);
        if constexpr (u2 == S4) {
            const int i = k - 1 - u1, j = l + 1 + u2;
            if (i >= 0 && j <= W - 1) {
                const double pij = sm.ringq[j & 7][i];
                if (pij > 0.) {
                    const int tij = pair_type(S[i], S[j]);
                    acc += pij * shape2(T, u1, u2, tij, t2, S[i + 1], S[j - 1], sp1, sq1) * sm.scale[u1 + u2 + 2];
                }
            }
        }
    });
    return acc;
}

__global__ void __launch_bounds__(NT2, 1)
pf2_kernel(PfLaunch L, const MfeTables *__restrict__ MT, const PfTables *__restrict__ T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem2 &sm = *reinterpret_cast<Smem2 *>(smem_raw);
    const int W = L.W;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned full = 0xffffffffu;
    double *qbG = L.gscratch + (long long)blockIdx.x * L.gscratch_per_cta;   // qb[j][i], pitch P2
    double *pmG = qbG + P2 * P2;                                              // PM[j][i]
    const int HF = (W + 3) / 2;
    auto qmidx = [&](int kk, int i) { return kk <= HF ? kk * PQ + i : (W + 3 - kk) * PQ + (W - kk) + i; };
    const double *ringp = &sm.ring[0][0][0];

    for (int k = tid; k < 200; k += NT2) {
        (&sm.th.expmismatchI[0][0][0])[k] = (&T->expmismatchI[0][0][0])[k];
        (&sm.th.expmismatch1nI[0][0][0])[k] = (&T->expmismatch1nI[0][0][0])[k];
        (&sm.th.expmismatchM[0][0][0])[k] = (&T->expmismatchM[0][0][0])[k];
        (&sm.th.expmismatchExt[0][0][0])[k] = (&T->expmismatchExt[0][0][0])[k];
        if (k < 40) {
            (&sm.th.expdangle5[0][0])[k] = (&T->expdangle5[0][0])[k];
            (&sm.th.expdangle3[0][0])[k] = (&T->expdangle3[0][0])[k];
        }
    }
    if (tid == 0) {
        sm.th.expMLintern = T->expMLintern;
        sm.th.expTermAU = T->expTermAU;
    }
    const PfHead *TH = &sm.th;
    if (tid == 0) {
        sm.scale[0] = 1.;
        sm.emlb[0] = 1.;
        sm.ainv[0] = 1.;
        const double x = T->expMLbase / T->pf_scale;
        for (int k = 1; k < P2 + 40; k++) sm.scale[k] = sm.scale[k - 1] / T->pf_scale;
        for (int k = 1; k < P2 + 8; k++) {
            sm.emlb[k] = sm.emlb[k - 1] * x;
            sm.ainv[k] = sm.ainv[k - 1] / x;
        }
    }
    __syncthreads();
    const double sc1 = sm.scale[1], sc2 = sm.scale[2], eml1 = sm.emlb[1];
    const double closing = T->expMLclosing, tAU = T->expTermAU;
    Ctx2 c;
    c.T = T;
    c.M = MT;
    c.S = sm.S;
    c.scale = sm.scale;
    c.W = W;
    const unsigned char *S = sm.S;
    auto nb = [&](int k) { return (k >= 0 && k < W) ? (int)S[k] : -1; };   // neighbour code or -1

    for (int fold = blockIdx.x; fold < L.n_fold; fold += gridDim.x) {
        __syncthreads();
        for (int k = tid; k < W; k += NT2) {
            sm.S[k] = L.seqs[(long long)fold * W + k];
            const char ch = L.hc ? (char)L.hc[(long long)fold * W + k] : '.';
            sm.can5[k] = !(ch == 'x' || ch == '>');
            sm.can3[k] = !(ch == 'x' || ch == '<');
        }
        for (int k = tid; k < P2 + 8; k += NT2) {
            sm.cen[k] = 0;
            sm.qm1[0][k] = 0.;
            sm.qm1[1][k] = 0.;
            sm.qqcol[k] = 0.;
            sm.x2[k] = 0.;
        }
        for (int k = tid; k < 3 * 32 * RP; k += NT2) (&sm.ring[0][0][0])[k] = 0.;
        for (int k = tid; k < 8 * P2; k += NT2) (&sm.ringq[0][0])[k] = 0.;
        if (tid == 0) {
            sm.q5[0] = 1.;
            for (int k = 1; k <= min(W, TURN + 1); k++) sm.q5[k] = sm.q5[k - 1] * sc1;
        }
        __syncthreads();

        // ================= inside, column j =================
        for (int j = TURN + 1; j < W; j++) {
            // ---- A: separable interior loops, warp = (32 cells) x (u2 group)
            {
                const int iblk = warp & 3, s = warp >> 2, i = iblk * 32 + lane;
                const int t = (i <= j - TURN - 1 && sm.can5[i] && sm.can3[j]) ? pair_type(S[i], S[j]) : 0;
                if (s == 0 && i < P2) sm.ty[i] = (unsigned char)t;
                if (__any_sync(full, t != 0)) {
                    if (t) {
                        double aG = 0., a1 = 0., aB = 0., aS = 0.;
                        const int u2max = j - 5;
                        switch (s) {
                            case 0: cand_group<0, false>(ringp, j - 1, i + 1, u2max, aG, a1, aB); aS = shapes_inside<0>(sm, T, i, j, t); break;
                            case 1: cand_group<1, false>(ringp, j - 1, i + 1, u2max, aG, a1, aB); aS = shapes_inside<1>(sm, T, i, j, t); break;
                            case 2: cand_group<2, false>(ringp, j - 1, i + 1, u2max, aG, a1, aB); aS = shapes_inside<2>(sm, T, i, j, t); break;
                            default: cand_group<3, false>(ringp, j - 1, i + 1, u2max, aG, a1, aB); aS = shapes_inside<3>(sm, T, i, j, t) + hairpin2(c, i, j, t) * sm.scale[j - i + 1]; break;
                        }
                        const int si1 = S[i + 1], sj1 = S[j - 1];
                        sm.partA[s][i] = aG * TH->expmismatchI[t][si1][sj1] + a1 * TH->expmismatch1nI[t][si1][sj1] +
                                         aB * (t > 2 ? tAU : 1.) + aS;
                    }
                }
            }
            __syncthreads();
            // ---- B: qb of the column, ring copies, qm1
            if (tid < RP) {
                const int i = tid;
                const int t = i <= j - TURN - 1 ? sm.ty[i] : 0;
                double qv = 0., vG = 0., v1 = 0., vB = 0.;
                if (t) {
                    const int si1 = S[i + 1], sj1 = S[j - 1];
                    qv = sm.partA[0][i] + sm.partA[1][i] + sm.partA[2][i] + sm.partA[3][i];
                    qv += sm.qqcol[i + 1] * closing * mlstem2(TH, rtype_of(t), sj1, si1) * sc2;
                    if (i > 0 && j < W - 1) {   // (i,j) as the inner pair of an enclosing loop
                        const int t2 = rtype_of(t), a = S[j + 1], b = S[i - 1];
                        vG = qv * TH->expmismatchI[t2][a][b];
                        v1 = qv * TH->expmismatch1nI[t2][a][b];
                        vB = t2 > 2 ? qv * tAU : qv;
                    }
                }
                const int slot = j & 31;
                sm.ring[0][slot][i] = vG;
                sm.ring[1][slot][i] = v1;
                sm.ring[2][slot][i] = vB;
                if (i < P2) {
                    sm.ringq[j & 7][i] = qv;
                    if (i < W) qbG[j * P2 + i] = qv;
                    double m1 = 0.;
                    if (i <= j - TURN - 1) {
                        if (j - 1 - i > TURN) m1 = sm.qm1[(j - 1) & 1][i] * eml1;
                        if (t) m1 += qv * mlstem2(TH, t, nb(i - 1), nb(j + 1));
                    }
                    sm.qm1[j & 1][i] = m1;
                }
            }
            __syncthreads();
            // ---- C: multiloop sums of the column (12 warps), the geometric part as a scan (warp 12), q5 (warp 13)
            const double *qm1c = sm.qm1[j & 1];
            if (warp < 12) {
                const int iblk = warp & 3, kg = warp >> 2, i = iblk * 32 + lane;
                const int n = j - 8;   // k' = 4 .. j-5
                double acc = 0.;
                if (n > 0) {
                    const int k0 = 4 + n * kg / 3, k1 = 4 + n * (kg + 1) / 3;
                    for (int kk = max(k0, iblk * 32 + 4); kk < k1; kk++)
                        if (i <= kk - 4) acc = fma(sm.qm[qmidx(kk, i)], qm1c[kk + 1], acc);
                }
                if (i < P2) sm.partC[kg][i] = acc;
            } else if (warp == 12) {
                // E[i] = sum_{k=i+1}^{j-4} eMLb[k-i] qm1[k,j] = ainv[i] * (suffix sum of eMLb[k] qm1[k,j])
                double tk[4], tot = 0.;
#pragma unroll
                for (int z = 0; z < 4; z++) {
                    const int k = 4 * lane + z;
                    tk[z] = (k <= j - TURN - 1 && k < P2) ? sm.emlb[k] * qm1c[k] : 0.;
                    tot += tk[z];
                }
                double inc = tot;
                for (int o = 1; o < 32; o <<= 1) {
                    const double u = __shfl_down_sync(full, inc, o);
                    if (lane + o < 32) inc += u;
                }
                double run = __shfl_down_sync(full, inc, 1);
                if (lane == 31) run = 0.;
#pragma unroll
                for (int z = 3; z >= 0; z--) {
                    const int k = 4 * lane + z;
                    if (k < P2) sm.ecol[k] = run * sm.ainv[k];
                    run += tk[z];
                }
            } else if (warp == 13) {
                double acc = 0.;
                for (int i = lane; i <= j - TURN - 1; i += 32) {
                    const int t = sm.ty[i];
                    if (t) acc += sm.q5[i] * sm.ringq[j & 7][i] * extloop2(TH, t, nb(i - 1), nb(j + 1));
                }
                acc = warp_sum(acc);
                if (lane == 0) sm.q5[j + 1] = sm.q5[j] * sc1 + acc;
            }
            __syncthreads();
            // ---- D: qm of the column (no barrier needed before the next column's phase A)
            if (tid <= j - TURN - 1) {
                const int i = tid;
                const double qq = sm.partC[0][i] + sm.partC[1][i] + sm.partC[2][i];
                sm.qqcol[i] = qq;
                sm.qm[qmidx(j, i)] = qm1c[i] + sm.ecol[i] + qq;
            } else if (tid < P2 + 8) {
                sm.qqcol[tid] = 0.;
            }
        }
        __syncthreads();
        const double Z = sm.q5[W];

        // ================= outside, column l =================
        for (int k = tid; k < 3 * 32 * RP; k += NT2) (&sm.ring[0][0][0])[k] = 0.;
        for (int k = tid; k < 8 * P2; k += NT2) (&sm.ringq[0][0])[k] = 0.;
        if (tid == 0) {
            sm.q3[W] = 1.;
            for (int k = W - 1; k >= max(0, W - TURN - 1); k--) sm.q3[k] = sm.q3[k + 1] * sc1;
        }
        double ed_local = 0.;
        __syncthreads();
        for (int l = W - 1; l > TURN; l--) {
            // ---- A: separable interior loops closed outside (k,l); partial X1; q3[l]
            {
                const int kblk = warp & 3, s = warp >> 2, k = kblk * 32 + lane;
                const double qkl = k <= l - TURN - 1 ? qbG[l * P2 + k] : 0.;
                const bool act = qkl != 0. && k >= 1 && l <= W - 2;
                if (s == 0 && k < P2) sm.qcol[k] = qkl;
                if (__any_sync(full, act)) {
                    if (act) {
                        double aG = 0., a1 = 0., aB = 0., aS = 0.;
                        const int u2max = W - 2 - l;
                        const int t2 = rtype_of(pair_type(S[k], S[l])), a = S[l + 1], b = S[k - 1];
                        switch (s) {
                            case 0: cand_group<0, true>(ringp, l + 1, 32 + k - 1, u2max, aG, a1, aB); aS = shapes_outside<0>(sm, T, k, l, t2, W); break;
                            case 1: cand_group<1, true>(ringp, l + 1, 32 + k - 1, u2max, aG, a1, aB); aS = shapes_outside<1>(sm, T, k, l, t2, W); break;
                            case 2: cand_group<2, true>(ringp, l + 1, 32 + k - 1, u2max, aG, a1, aB); aS = shapes_outside<2>(sm, T, k, l, t2, W); break;
                            default: cand_group<3, true>(ringp, l + 1, 32 + k - 1, u2max, aG, a1, aB); aS = shapes_outside<3>(sm, T, k, l, t2, W); break;
                        }
                        sm.partA[s][k] = aG * TH->expmismatchI[t2][a][b] + a1 * TH->expmismatch1nI[t2][a][b] +
                                         aB * (t2 > 2 ? tAU : 1.) + aS;
                    }
                }
                // X1[i,l] = sum_{j >= l+6} PM[i,j] qm[l+1,j-1], the j range cut in four
                const int i = k, nj = W - l - 6;
                double acc = 0.;
                if (nj > 0 && i <= l - TURN - 1) {
                    // PM streams from L2: four independent loads in flight per lane instead of one per trip
                    const int j0 = l + 6 + nj * s / 4, j1 = l + 6 + nj * (s + 1) / 4;
                    double acc2 = 0.;
                    int j = j0;
                    for (; j + 3 < j1; j += 4) {
                        const double p0 = pmG[j * P2 + i], p1 = pmG[(j + 1) * P2 + i], p2 = pmG[(j + 2) * P2 + i], p3 = pmG[(j + 3) * P2 + i];
                        acc = fma(p0, sm.qm[qmidx(j - 1, l + 1)], acc);
                        acc2 = fma(p1, sm.qm[qmidx(j, l + 1)], acc2);
                        acc = fma(p2, sm.qm[qmidx(j + 1, l + 1)], acc);
                        acc2 = fma(p3, sm.qm[qmidx(j + 2, l + 1)], acc2);
                    }
                    for (; j < j1; j++) acc = fma(pmG[j * P2 + i], sm.qm[qmidx(j - 1, l + 1)], acc);
                    acc += acc2;
                }
                if (i < P2) sm.partC[s][i] = acc;
                if (warp == 15) {   // q3[l] for the next column
                    double a3 = 0.;
                    for (int j = l + TURN + 1 + lane; j < W; j += 32) {
                        const double q = qbG[j * P2 + l];
                        if (q != 0.) a3 += q * sm.q3[j + 1] * extloop2(TH, pair_type(S[l], S[j]), nb(l - 1), nb(j + 1));
                    }
                    a3 = warp_sum(a3);
                    if (lane == 0) sm.q3[l] = sm.q3[l + 1] * sc1 + a3;
                }
            }
            __syncthreads();
            // ---- B: X1, X2 of the column; geometric part of the multiloop term as a prefix scan (warp 12)
            if (tid < P2) {
                const int i = tid;
                double x1 = 0., x2 = 0.;
                if (i <= l - TURN - 1) {
                    x1 = sm.partC[0][i] + sm.partC[1][i] + sm.partC[2][i] + sm.partC[3][i];
                    if (l + 1 < W) x2 = sm.x2[i] * eml1 + (i <= l - TURN ? sm.pmcol[i] : 0.);
                }
                sm.x1[i] = x1;
                sm.x2[i] = x2;
                sm.x12[i] = x1 + x2;
            } else if (warp == 12) {
                // G1[k] = sum_{i<k} X1[i] eMLb[k-1-i] = eMLb[k-1] * (prefix sum of X1[i] ainv[i])
                double tk[4], tot = 0.;
#pragma unroll
                for (int z = 0; z < 4; z++) {
                    const int i = 4 * lane + z;
                    double x1 = 0.;
                    if (i <= l - TURN - 1 && i < P2)
                        x1 = sm.partC[0][i] + sm.partC[1][i] + sm.partC[2][i] + sm.partC[3][i];
                    tk[z] = x1 * sm.ainv[min(i, P2 + 7)];
                    tot += tk[z];
                }
                double inc = tot;
                for (int o = 1; o < 32; o <<= 1) {
                    const double u = __shfl_up_sync(full, inc, o);
                    if (lane >= o) inc += u;
                }
                double run = __shfl_up_sync(full, inc, 1);
                if (lane == 0) run = 0.;
#pragma unroll
                for (int z = 0; z < 4; z++) {
                    const int k = 4 * lane + z;
                    if (k < P2) sm.g1[k] = k >= 1 ? run * sm.emlb[k - 1] : 0.;
                    run += tk[z];
                }
            }
            __syncthreads();
            // ---- C: H[k] = sum_{i <= k-6} (X1+X2)[i] qm[i+1,k-1], the i range cut in three
            if (warp < 12) {
                const int kblk = warp & 3, g = warp >> 2, k = kblk * 32 + lane;
                const int n = l - 9;   // i = 0 .. l-10
                double acc = 0.;
                if (n > 0 && k <= l - TURN - 1) {
                    const int i0 = n * g / 3, i1 = n * (g + 1) / 3;
                    for (int i = i0; i < i1; i++)
                        if (i <= k - 6) acc = fma(sm.x12[i], sm.qm[qmidx(k - 1, i + 1)], acc);
                }
                if (k < P2) sm.partC[g][k] = acc;
            }
            __syncthreads();
            // ---- D: P of the column, ring copies, PM, probabilities
            if (tid < RP - 32) {
                const int k = tid;
                double Pv = 0., vG = 0., v1 = 0., vB = 0., pm = 0.;
                if (k <= l - TURN - 1) {
                    const double qkl = sm.qcol[k];
                    if (qkl != 0.) {
                        const int t = pair_type(S[k], S[l]);
                        if (k >= 1 && l <= W - 2) {
                            const int t2 = rtype_of(t), sp1 = S[k - 1], sq1 = S[l + 1];
                            Pv = sm.partA[0][k] + sm.partA[1][k] + sm.partA[2][k] + sm.partA[3][k];
                            const double ml = sm.g1[k] + sm.partC[0][k] + sm.partC[1][k] + sm.partC[2][k];
                            Pv += ml * mlstem2(TH, t, sp1, sq1) * sc2;
                        }
                        Pv += sm.q5[k] * sm.q3[l + 1] / Z * extloop2(TH, t, nb(k - 1), nb(l + 1));
                        if (Pv != 0.) {
                            const int a = S[k + 1], b = S[l - 1];
                            vG = Pv * TH->expmismatchI[t][a][b];
                            v1 = Pv * TH->expmismatch1nI[t][a][b];
                            vB = t > 2 ? Pv * tAU : Pv;
                        }
                        pm = Pv * closing * mlstem2(TH, rtype_of(t), S[l - 1], S[k + 1]);
                        const double p = Pv * qkl;
                        ed_local += p * (1. - p);
                        if (p > 0.5) {
                            sm.cen[k] = (short)(l + 1);
                            sm.cen[l] = (short)(k + 1);
                        }
                        if (L.bpp) L.bpp[((long long)fold * W + k) * W + l] = p;
                    }
                    pmG[l * P2 + k] = pm;
                    sm.pmcol[k] = pm;
                }
                const int slot = l & 31;
                sm.ring[0][slot][32 + k] = vG;
                sm.ring[1][slot][32 + k] = v1;
                sm.ring[2][slot][32 + k] = vB;
                sm.ringq[l & 7][k] = Pv;
            }
            __syncthreads();
        }

        // ================= ED, centroid =================
        ed_local = warp_sum(ed_local);
        if (lane == 0) sm.red[warp] = ed_local;
        __syncthreads();
        if (tid == 0) {
            double s = 0.;
            for (int w = 0; w < NT2 / 32; w++) s += sm.red[w];
            L.ed[fold] = 2. * s;
            L.dG[fold] = (-log(Z) - W * log(T->pf_scale)) * T->kT / 1000.;
        }
        for (int k = tid; k < W; k += NT2) L.centroid[(long long)fold * W + k] = sm.cen[k];
    }
}

}  // namespace

static bool &pf2_enabled() {
    static bool on = !(getenv("SFB_PF_ENGINE") && atoi(getenv("SFB_PF_ENGINE")) == 1);
    return on;
}
void pf2_set_enabled(bool on) { pf2_enabled() = on; }

bool pf2_supports(const PfLaunch &L) {
    return pf2_enabled() && (!L.hc || L.hc_simple) && !L.sc && L.max_span <= 0 && L.W >= 2 * TURN + 4 && L.W <= P2;
}

size_t pf2_scratch_doubles_per_cta() { return 2 * (size_t)P2 * P2; }

void pf2_upload_tables(const PfTables &q) {
    static double g[32 * 32], o[32], b[32];
    double scale[40];
    scale[0] = 1.;
    for (int k = 1; k < 40; k++) scale[k] = scale[k - 1] / q.pf_scale;
    for (int u1 = 0; u1 < 32; u1++)
        for (int u2 = 0; u2 < 32; u2++) {
            const int u = u1 + u2, d = u1 > u2 ? u1 - u2 : u2 - u1;
            g[u1 * 32 + u2] = u <= MAXLOOP ? q.expinternal[u] * q.expninio[d] * scale[u + 2] : 0.;
        }
    for (int u = 0; u < 32; u++) {
        o[u] = (u >= 2 && u <= MAXLOOP) ? q.expinternal[u] * q.expninio[u - 2] * scale[u + 2] : 0.;
        b[u] = (u >= 1 && u <= MAXLOOP) ? q.expbulge[u] * scale[u + 2] : 0.;
    }
    cudaMemcpyToSymbol(c2_G, g, sizeof(g));
    cudaMemcpyToSymbol(c2_1, o, sizeof(o));
    cudaMemcpyToSymbol(c2_B, b, sizeof(b));
}

void launch_pf2(const PfLaunch &L, const MfeTables *d_mfe, const PfTables *d_pf, int n_sm, cudaStream_t stream,
                int *n_launches) {
    const size_t smem = sizeof(Smem2);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(pf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    const int grid = L.n_fold < n_sm ? L.n_fold : n_sm;
    pf2_kernel<<<grid, NT2, smem, stream>>>(L, d_mfe, d_pf);
    if (n_launches) (*n_launches)++;
}

}  // namespace sfb
