// Partition function, second generation: one CTA (512 threads) per window, everything the inner loops touch in
// shared memory, windows up to 120 nt without constraints or with per-nucleotide hard constraints ('x' '<' '>').
//
// Replaces fc.pf(), fc.centroid(), fc.mean_bp_distance() -- ScanFold.py:498,503-504 -- for the native windows of a
// scan; pf.cu keeps enforced pairs, soft constraints and the long windows.
//
// Both passes walk the matrix by COLUMN (3' end): every cell of a column only depends on earlier columns.  r02
// rewrite (the first version spent 6,000 cycles per column in a lane = cell candidate walk whose shared-memory loads
// served one useful lane in three, 3,000 in a serial multiloop sum and three to four barriers per column):
// * separable interior loops: ONE WARP PER PAIRABLE CELL (a compacted list per column), lane = u2, a compile-time
//   loop over u1.  The candidates' weights -- qb (inside) / the outside weight P (outside), already multiplied by the
//   mismatch factor of that pair for the three separable loop classes (generic | 1xn | bulge) -- live in rings
//   [position][column & 31] with a pitch of 33 doubles: the 32 lanes of a load read 32 consecutive doubles (no bank
//   conflict), the u1 step is an immediate offset (no address arithmetic), the size factor K[u1] of lane u2 sits in a
//   register: one LDS.64 + one DFMA per 32 candidates, then one warp reduction per cell.  Lanes 0 and 1 (u2 = 0, 1) read
//   the bulge / 1xn copy, every other lane the generic one; u1 = 0, 1 are peeled the same way.
// * everything else of a column runs beside that walk or in ONE second phase as warp-sized units (two barriers per
//   column): the multiloop sums as row-sliced matrix-vector products with four independent accumulators per lane,
//   the nine table-driven shapes and the hairpin of the NEXT column (lane = cell, all table loads issued together),
//   the geometric sums as warp scans, the exterior sums, the ring bookkeeping.
// qm lives in a folded triangular matrix with an odd pitch (row and column walks are both bank-conflict free).
// Only qb (read back once per cell) and the multiloop closing weights PM (read as rows) stream through L2.
#include <cstdlib>
#include <type_traits>

#include "device_common.cuh"

namespace sfb {
namespace {

constexpr int P2 = 120;          // longest window
constexpr int PT = 33;           // ring pitch (doubles): [position][column & 31] plus one pad
constexpr int RPAD = 4;          // zero positions either side: a group of four u1 steps may overrun by three
constexpr int RPOS = P2 + 2 * RPAD;
constexpr int RING = RPOS * PT;  // doubles per ring copy
constexpr int PQ = 121;          // pitch of the folded qm matrix (odd)
constexpr int QROWS = (P2 + 3) / 2 + 1;
constexpr int NT2 = 512, NW2 = NT2 / 32;
constexpr int NG = 8;            // inside: partial sums of the multiloop product (one warp each)
constexpr int NH = 7;            // outside: partial sums of H
constexpr int NX = 3;            // outside: partial sums of X1
constexpr int PP = 128;          // pitch of the per-cell arrays

__device__ double g_K2[32 * 32];  // [u2][u1]: size factor of the separable candidate (u1, u2), 0 for the table-driven ones

// the small Boltzmann-factor tables the per-cell code looks up (same member names as PfTables): a shared-memory copy
struct PfHead {
    double expmismatchI[8][5][5], expmismatch1nI[8][5][5], expmismatchM[8][5][5], expmismatchExt[8][5][5];
    double expdangle5[8][5], expdangle3[8][5];
    double expMLintern, expTermAU;
};

struct Smem2 {
    PfHead th;
    double ring[3 * RING];      // generic | 1xn | bulge copies, [position + RPAD][column & 31]
    double ringq[8][P2];        // raw qb (inside) / raw P (outside) of the last columns: table-driven shapes
    double qm[QROWS * PQ];      // folded: row = 3' end k', entries i <= k'-4
    double partC[NG][PP];       // partial multiloop sums (inside: qq; outside: H)
    double partX[NX][PP];       // outside: partial X1 of the next column
    double partA[PP];           // outside: interior-loop sum of the listed cells
    double partS[PP];           // table-driven shapes (+ hairpin) of the coming column
    double qm1[2][PP];
    double ecol[PP];
    double x1[PP], x2[PP], x12[PP], g1[PP];
    double q5[PP], q3[PP];
    double qcol[PP];            // outside: qb of the current column
    double scale[P2 + 40], emlb[PP], ainv[PP];
    double red[32];
    short cen[PP];
    unsigned char S[PP];
    unsigned char ty[2][PP];    // pair type of the cells of a column (0: not pairable)
    unsigned char list[2][PP];  // 5' ends of the cells the candidate walk visits
    unsigned char can5[PP], can3[PP];   // hard constraints: may be the 5' / 3' partner of a pair
    int cnt[2];
#ifdef SFB_PF2_TIMING
    long long tstamp[2][16];
#endif
};

static_assert(sizeof(Smem2) <= 227 * 1024, "one window per SM");

template <int A, int B, class F>
__device__ __forceinline__ void sfor2(F &&f) {
    if constexpr (A <= B) {
        f(std::integral_constant<int, A>{});
        sfor2<A + 1, B>(f);
    }
}

// Separable candidates of one cell: lane = u2, u1 = 0 .. u1max.  pB / p1 / pM point at position u1 = 0 of the bulge /
// 1xn / lane-class copy in this lane's column; inside steps up the positions, outside down.  Returns the lane-class sum
// (u1 >= 2); accB (u1 = 0: bulges) and acc1 (u1 = 1: 1xn loops) take their own outer factor.
template <bool OUT>
__device__ __forceinline__ double cand_walk(const double *pB, const double *p1, const double *pM, int u1max,
                                            const double (&K)[MAXLOOP + 1], double &accB, double &acc1) {
    constexpr int ST = OUT ? -PT : PT;
    accB = pB[0] * K[0];
    acc1 = p1[ST] * K[1];
    double a0 = 0., a1 = 0.;
    sfor2<0, 7>([&](auto Q) {
        constexpr int g = 2 + 4 * decltype(Q)::value;
        if (g <= u1max) {   // warp-uniform
            sfor2<0, 3>([&](auto V) {
                constexpr int u1 = g + decltype(V)::value;
                if constexpr (u1 <= MAXLOOP) {
                    if constexpr (u1 & 1)
                        a1 = fma(pM[u1 * ST], K[u1], a1);
                    else
                        a0 = fma(pM[u1 * ST], K[u1], a0);
                }
            });
        }
    });
    return a0 + a1;
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

// the nine table-driven shapes closed by (i,j) plus the hairpin, lane = cell; every load is unconditional (a cell that
// is no pair carries weight 0), so the table loads of all shapes are in flight together
__device__ __forceinline__ double shapes_inside(const Smem2 &sm, const PfTables *T, const Ctx2 &c, int i, int j, int t) {
    const unsigned char *S = sm.S;
    const int si1 = S[i + 1], sj1 = S[j - 1];
    double acc = 0.;
    sfor2<0, 8>([&](auto Z) {
        constexpr int z = decltype(Z)::value, u1 = shape_u1(z), u2 = shape_u2(z);
        const int p = i + 1 + u1, q = j - 1 - u2;
        const double qpq = sm.ringq[q & 7][p];
        const int t2 = rtype_of(pair_type(S[p], S[q]));
        const double f = shape2(T, u1, u2, t, t2, si1, sj1, S[p - 1], S[q + 1]) * sm.scale[u1 + u2 + 2];
        acc += qpq != 0. ? qpq * f : 0.;
    });
    return acc + hairpin2(c, i, j, t) * sm.scale[j - i + 1];
}

// the same for the outside pass: (k,l) is the inner pair, (i,j) = (k-1-u1, l+1+u2) the closing one
__device__ __forceinline__ double shapes_outside(const Smem2 &sm, const PfTables *T, int k, int l, int t2, int W) {
    const unsigned char *S = sm.S;
    const int sp1 = S[k - 1], sq1 = S[l + 1];
    double acc = 0.;
    sfor2<0, 8>([&](auto Z) {
        constexpr int z = decltype(Z)::value, u1 = shape_u1(z), u2 = shape_u2(z);
        const int i = k - 1 - u1, j = l + 1 + u2;
        const bool ok = i >= 0 && j <= W - 1;
        const int ic = max(i, 0), jc = min(j, W - 1);
        const double pij = ok ? sm.ringq[jc & 7][ic] : 0.;
        const int tij = pair_type(S[ic], S[jc]);
        const double f = shape2(T, u1, u2, tij, t2, S[ic + 1], S[jc - 1], sp1, sq1) * sm.scale[u1 + u2 + 2];
        acc += pij > 0. ? pij * f : 0.;
    });
    return acc;
}

#ifdef SFB_PF2_TIMING
// barrier with bookkeeping: work[slot] += own arrival - release of the previous barrier, wait[slot] += last arrival - own
#define PF2_SYNC(slot)                                                                     \
    {                                                                                      \
        long long now_;                                                                    \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(now_)::"memory");                     \
        if (lane == 0) sm.tstamp[tpar][warp] = now_;                                       \
        __syncthreads();                                                                   \
        long long m_ = sm.tstamp[tpar][lane & 15];                                         \
        for (int o_ = 8; o_; o_ >>= 1) m_ = max(m_, __shfl_xor_sync(0xffffffffu, m_, o_)); \
        tacc[slot] += now_ - tlast;                                                        \
        tacc[slot + 1] += m_ - now_;                                                       \
        tlast = m_;                                                                        \
        tpar ^= 1;                                                                         \
    }
#define PF2_RESET asm volatile("mov.u64 %0, %%clock64;" : "=l"(tlast)::"memory");
#define PF2_PROBE(k)                                                    \
    {                                                                   \
        long long now_;                                                 \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(now_)::"memory"); \
        tprobe[k] += now_ - tp_;                                        \
        tp_ = now_;                                                     \
    }
#define PF2_PROBE0 asm volatile("mov.u64 %0, %%clock64;" : "=l"(tp_)::"memory");
#else
#define PF2_PROBE(k)
#define PF2_PROBE0
#define PF2_SYNC(slot) __syncthreads();
#define PF2_RESET
#endif

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
    auto qmrow = [&](int kk) { return kk <= HF ? kk * PQ : (W + 3 - kk) * PQ + (W - kk); };   // + i
#ifdef SFB_PF2_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast;
    int tpar = 0;
    long long tprobe[4] = {0, 0, 0, 0}, tp_ = 0;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(tlast)::"memory");
#endif

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
        for (int k = 1; k < PP; k++) {
            sm.emlb[k] = sm.emlb[k - 1] * x;
            sm.ainv[k] = sm.ainv[k - 1] / x;
        }
    }
    double K[MAXLOOP + 1];   // size factors of this lane's candidates (lane = u2)
#pragma unroll
    for (int u1 = 0; u1 <= MAXLOOP; u1++) K[u1] = g_K2[lane * 32 + u1];
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
    const int cls = lane == 0 ? 2 : (lane == 1 ? 1 : 0);   // ring copy of this lane's candidates with u1 >= 2

    // pairable cells of inside column j -> ty / list / cnt of parity j & 1 (one warp)
    auto build_list_inside = [&](int j) {
        int n = 0;
        const bool c3 = j < W && sm.can3[j];
        for (int b = 0; b < PP / 32; b++) {
            const int i = b * 32 + lane;
            const int t = (c3 && i <= j - TURN - 1 && sm.can5[i]) ? pair_type(S[i], S[j]) : 0;
            sm.ty[j & 1][i] = (unsigned char)t;
            const unsigned m = __ballot_sync(full, t != 0);
            if (t) sm.list[j & 1][n + __popc(m & ((1u << lane) - 1))] = (unsigned char)i;
            n += __popc(m);
        }
        if (lane == 0) sm.cnt[j & 1] = n;
    };
    // table-driven shapes + hairpin of inside column j, cells 32 b .. 32 b + 31 (one warp)
    auto shapes_unit_inside = [&](int j, int b) {
        const int i = b * 32 + lane;
        double v = 0.;
        if (j < W && i <= j - TURN - 1 && sm.can5[i] && sm.can3[j]) {
            const int t = pair_type(S[i], S[j]);
            if (t) v = shapes_inside(sm, T, c, i, j, t);
        }
        sm.partS[i] = v;
    };

    for (int fold = blockIdx.x; fold < L.n_fold; fold += gridDim.x) {
        __syncthreads();
        for (int k = tid; k < PP; k += NT2) {
            const bool in = k < W;
            sm.S[k] = in ? L.seqs[(long long)fold * W + k] : 4;
            const char ch = (L.hc && in) ? (char)L.hc[(long long)fold * W + k] : '.';
            sm.can5[k] = in && !(ch == 'x' || ch == '>');
            sm.can3[k] = in && !(ch == 'x' || ch == '<');
            sm.cen[k] = 0;
            sm.qm1[0][k] = 0.;
            sm.qm1[1][k] = 0.;
            sm.ecol[k] = 0.;
            sm.x1[k] = 0.;
            sm.x2[k] = 0.;
            sm.x12[k] = 0.;
            sm.g1[k] = 0.;
            sm.partA[k] = 0.;
            sm.partS[k] = 0.;
        }
        for (int k = tid; k < 3 * RING; k += NT2) sm.ring[k] = 0.;
        for (int k = tid; k < 8 * P2; k += NT2) (&sm.ringq[0][0])[k] = 0.;
        for (int k = tid; k < NG * PP; k += NT2) (&sm.partC[0][0])[k] = 0.;
        for (int k = tid; k < NX * PP; k += NT2) (&sm.partX[0][0])[k] = 0.;
        if (tid == 0) {
            sm.q5[0] = 1.;
            for (int k = 1; k <= min(W, TURN + 1); k++) sm.q5[k] = sm.q5[k - 1] * sc1;
        }
        __syncthreads();
        if (warp == 15) build_list_inside(TURN + 1);
        if (warp >= 8 && warp < 12) shapes_unit_inside(TURN + 1, warp - 8);
        __syncthreads();

        PF2_RESET
        // ================= inside, column j =================
        for (int j = TURN + 1; j < W; j++) {
            const int par = j & 1;
            // ---- phase 1: the candidate walk of the pairable cells; beside it (warps 0-3, lane = cell) the cells that
            // are no pair and qm of the previous column
            if (warp < 4) {
                const int i = tid;
                if (!sm.ty[par][i] && i < P2) {
                    sm.ringq[j & 7][i] = 0.;
                    if (i < W) qbG[j * P2 + i] = 0.;
                    sm.qm1[par][i] = i <= j - TURN - 2 ? sm.qm1[par ^ 1][i] * eml1 : 0.;
                }
                if (j - 1 > TURN && i <= j - TURN - 2) {
                    double qq = 0.;
#pragma unroll
                    for (int g = 0; g < NG; g++) qq += sm.partC[g][i];
                    sm.qm[qmrow(j - 1) + i] = sm.qm1[par ^ 1][i] + sm.ecol[i] + qq;
                }
            }
            {
                const int n = sm.cnt[par];
                const double *qm1p = sm.qm1[par ^ 1];
                PF2_PROBE0
                for (int cc = (warp + NW2 - 4) & (NW2 - 1); cc < n; cc += NW2) {
                    PF2_PROBE(3)
                    const int i = sm.list[par][cc];
                    const int t = sm.ty[par][i];
                    const int si1 = S[i + 1], sj1 = S[j - 1];
                    const double mmI = TH->expmismatchI[t][si1][sj1], mm1 = TH->expmismatch1nI[t][si1][sj1];
                    const double tau = t > 2 ? tAU : 1.;
                    const double *pb = sm.ring + (i + 1 + RPAD) * PT + ((j - 1 - lane) & 31);
                    double accB, acc1;
                    const double aM = cand_walk<false>(pb + 2 * RING, pb + RING, pb + cls * RING, min(MAXLOOP, j - i - 6), K, accB, acc1);
                    PF2_PROBE(0)
                    double v = aM * (lane == 0 ? tau : (lane == 1 ? mm1 : mmI)) + accB * tau + acc1 * mm1;
                    if (lane < NG)   // multiloop closed by (i,j): the partial sums of the previous column join the reduction
                        v = fma(sm.partC[lane][i + 1], closing * mlstem2(TH, rtype_of(t), sj1, si1) * sc2, v);
                    else if (lane == NG)
                        v += sm.partS[i];
                    v = warp_sum(v);
                    PF2_PROBE(1)
                    // qb of the cell is known to every lane: ring copies, raw copy, qm1
                    if (lane < 3) {
                        double m = 0.;
                        if (i > 0 && j < W - 1) {   // (i,j) as the inner pair of an enclosing loop
                            const int t2 = rtype_of(t), a = S[j + 1], b = S[i - 1];
                            m = lane == 0 ? TH->expmismatchI[t2][a][b] : (lane == 1 ? TH->expmismatch1nI[t2][a][b] : (t2 > 2 ? tAU : 1.));
                        }
                        sm.ring[lane * RING + (i + RPAD) * PT + (j & 31)] = v * m;
                    } else if (lane == 3) {
                        sm.ringq[j & 7][i] = v;
                        qbG[j * P2 + i] = v;
                    } else if (lane == 4) {
                        const double m1 = i <= j - TURN - 2 ? qm1p[i] * eml1 : 0.;
                        sm.qm1[par][i] = m1 + v * mlstem2(TH, t, nb(i - 1), nb(j + 1));
                    }
                    PF2_PROBE(2)
                }
            }
            PF2_SYNC(0)
            // ---- phase 2: one unit per warp
            const double *qm1c = sm.qm1[par];
            if (warp < NG) {
                // qq[i] = sum_k' qm[i,k'-1] qm1[k',j]: rows k' = 4 + warp, + NG, ..; lane = cell in four blocks
                double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
#pragma unroll 2
                for (int kk = 4 + warp; kk <= j - 5; kk += NG) {
                    const double w = qm1c[kk + 1];
                    const double *row = sm.qm + qmrow(kk) + lane;
                    const int lim = kk - 4 - lane;   // cell i = 32 b + lane takes part if 32 b <= lim
                    if (lim >= 0) a0 = fma(row[0], w, a0);
                    if (lim >= 32) a1 = fma(row[32], w, a1);
                    if (lim >= 64) a2 = fma(row[64], w, a2);
                    if (lim >= 96) a3 = fma(row[96], w, a3);
                }
                sm.partC[warp][lane] = a0;
                sm.partC[warp][lane + 32] = a1;
                sm.partC[warp][lane + 64] = a2;
                sm.partC[warp][lane + 96] = a3;
            } else if (warp < NG + 4) {
                shapes_unit_inside(j + 1, warp - NG);
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
                    const int t = sm.ty[par][i];
                    if (t) acc += sm.q5[i] * sm.ringq[j & 7][i] * extloop2(TH, t, nb(i - 1), nb(j + 1));
                }
                acc = warp_sum(acc);
                if (lane == 0) sm.q5[j + 1] = sm.q5[j] * sc1 + acc;
            } else if (warp == 14) {
                // the ring column the next column writes: its last reader (column j) is done
                const int slot = (j + 1) & 31;
                for (int p = lane; p < RPOS; p += 32) {
                    sm.ring[p * PT + slot] = 0.;
                    sm.ring[RING + p * PT + slot] = 0.;
                    sm.ring[2 * RING + p * PT + slot] = 0.;
                }
            } else {
                build_list_inside(j + 1);
            }
            PF2_SYNC(2)
        }
        const double Z = sm.q5[W];

        // ================= outside, column l =================
        for (int k = tid; k < 3 * RING; k += NT2) sm.ring[k] = 0.;
        for (int k = tid; k < 8 * P2; k += NT2) (&sm.ringq[0][0])[k] = 0.;
        for (int k = tid; k < NG * PP; k += NT2) (&sm.partC[0][0])[k] = 0.;
        if (tid < PP) {
            // qm of the last column is never read; qb of the first outside column
            sm.qcol[tid] = (tid <= W - 1 - TURN - 1) ? qbG[(W - 1) * P2 + tid] : 0.;
            sm.partS[tid] = 0.;
        }
        if (tid == 0) {
            sm.q3[W] = 1.;
            for (int k = W - 1; k >= max(0, W - TURN - 1); k--) sm.q3[k] = sm.q3[k + 1] * sc1;
            sm.cnt[(W - 1) & 1] = 0;   // cells of the last column have no enclosing pair
        }
        double ed_local = 0.;
        __syncthreads();
        PF2_RESET
        for (int l = W - 1; l > TURN; l--) {
            const int par = l & 1;
            // ---- phase 1: one unit per warp, then the candidate walk of the listed cells
            if (warp < 4) {
                // table-driven shapes closed outside (k,l)
                const int k = warp * 32 + lane;
                double v = 0.;
                if (l <= W - 2 && k >= 1 && k <= l - TURN - 1 && sm.qcol[k] != 0.)
                    v = shapes_outside(sm, T, k, l, rtype_of(pair_type(S[k], S[l])), W);
                sm.partS[k] = v;
            } else if (warp < 4 + NH) {
                // H[k] = sum_{i <= k-6} (X1+X2)[i] qm[i+1,k-1]: i = g, g + NH, ..; lane = cell in four blocks
                const int g = warp - 4;
                double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
                const double *r0 = sm.qm + qmrow(max(lane - 1, 0)) + 1, *r1 = sm.qm + qmrow(lane + 31) + 1;
                const double *r2 = sm.qm + qmrow(min(lane + 63, W)) + 1, *r3 = sm.qm + qmrow(min(lane + 95, W)) + 1;
                const int kmax = l - TURN - 1;   // cells k <= kmax
#pragma unroll 2
                for (int i = g; i <= l - 10; i += NH) {
                    const double w = sm.x12[i];
                    if (i <= lane - 6 && lane <= kmax) a0 = fma(r0[i], w, a0);
                    if (i <= lane + 26 && lane + 32 <= kmax) a1 = fma(r1[i], w, a1);
                    if (i <= lane + 58 && lane + 64 <= kmax) a2 = fma(r2[i], w, a2);
                    if (i <= lane + 90 && lane + 96 <= kmax) a3 = fma(r3[i], w, a3);
                }
                sm.partC[g][lane] = a0;
                sm.partC[g][lane + 32] = a1;
                sm.partC[g][lane + 64] = a2;
                sm.partC[g][lane + 96] = a3;
            } else if (warp < 4 + NH + NX) {
                // X1[i,l-1] = sum_{j >= l+5} PM[i,j] qm[l,j-1] of the NEXT column (PM of those columns is final): the j
                // range cut in NX, PM streams from L2 with four independent loads in flight per lane
                const int g = warp - 4 - NH, lx = l - 1;
                const int nj = W - lx - 6;
                double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
                if (nj > 0 && lx > TURN) {
                    const int j0 = lx + 6 + nj * g / NX, j1 = lx + 6 + nj * (g + 1) / NX;
                    const int imax = lx - TURN - 1;
#pragma unroll 2
                    for (int j = j0; j < j1; j++) {
                        const double w = sm.qm[qmrow(j - 1) + lx + 1];
                        const double *pr = pmG + j * P2 + lane;
                        if (lane <= imax) a0 = fma(pr[0], w, a0);
                        if (lane + 32 <= imax) a1 = fma(pr[32], w, a1);
                        if (lane + 64 <= imax) a2 = fma(pr[64], w, a2);
                        if (lane + 96 <= imax) a3 = fma(pr[96], w, a3);
                    }
                }
                sm.partX[g][lane] = a0;
                sm.partX[g][lane + 32] = a1;
                sm.partX[g][lane + 64] = a2;
                sm.partX[g][lane + 96] = a3;
            } else if (warp == 14) {
                // G1[k] = sum_{i<k} X1[i] eMLb[k-1-i] = eMLb[k-1] * (prefix sum of X1[i] ainv[i])
                double tk[4], tot = 0.;
#pragma unroll
                for (int z = 0; z < 4; z++) {
                    const int i = 4 * lane + z;
                    tk[z] = (i <= l - TURN - 1 && i < P2) ? sm.x1[i] * sm.ainv[i] : 0.;
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
            } else {
                // q3[l] for the next column
                double a3 = 0.;
                for (int j = l + TURN + 1 + lane; j < W; j += 32) {
                    const double q = qbG[j * P2 + l];
                    if (q != 0.) a3 += q * sm.q3[j + 1] * extloop2(TH, pair_type(S[l], S[j]), nb(l - 1), nb(j + 1));
                }
                a3 = warp_sum(a3);
                if (lane == 0) sm.q3[l] = sm.q3[l + 1] * sc1 + a3;
            }
            {
                const int n = sm.cnt[par];
                for (int cc = (warp + 2) & (NW2 - 1); cc < n; cc += NW2) {
                    const int k = sm.list[par][cc];
                    const int t2 = rtype_of(pair_type(S[k], S[l])), a = S[l + 1], b = S[k - 1];
                    const double mmI = TH->expmismatchI[t2][a][b], mm1 = TH->expmismatch1nI[t2][a][b];
                    const double tau = t2 > 2 ? tAU : 1.;
                    const double *pb = sm.ring + (k - 1 + RPAD) * PT + ((l + 1 + lane) & 31);
                    double accB, acc1;
                    const double aM = cand_walk<true>(pb + 2 * RING, pb + RING, pb + cls * RING, min(MAXLOOP, k - 1), K, accB, acc1);
                    double v = aM * (lane == 0 ? tau : (lane == 1 ? mm1 : mmI)) + accB * tau + acc1 * mm1;
                    v = warp_sum(v);
                    if (lane == 0) sm.partA[k] = v;
                }
            }
            PF2_SYNC(4)
            // ---- phase 2: P of the column, ring copies, PM, probabilities; X1, X2 and qb of the next column
            if (tid < PP) {
                const int k = tid;
                double Pv = 0., vG = 0., v1 = 0., vB = 0., pm = 0.;
                if (k <= l - TURN - 1) {
                    const double qkl = sm.qcol[k];
                    if (qkl != 0.) {
                        const int t = pair_type(S[k], S[l]);
                        if (k >= 1 && l <= W - 2) {
                            const int sp1 = S[k - 1], sq1 = S[l + 1];
                            Pv = sm.partA[k] + sm.partS[k];
                            double ml = sm.g1[k];
#pragma unroll
                            for (int g = 0; g < NH; g++) ml += sm.partC[g][k];
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
                }
                if (k < P2) {
                    const int o = (k + RPAD) * PT + (l & 31);
                    sm.ring[o] = vG;
                    sm.ring[RING + o] = v1;
                    sm.ring[2 * RING + o] = vB;
                    sm.ringq[l & 7][k] = Pv;
                }
                // the next column l-1: X1 from the partial sums, X2 by its recurrence, qb
                const int lx = l - 1;
                double x1 = 0., x2 = 0., qn = 0.;
                if (lx > TURN && k <= lx - TURN - 1) {
                    x1 = sm.partX[0][k] + sm.partX[1][k] + sm.partX[2][k];
                    x2 = sm.x2[k] * eml1 + pm;
                    qn = qbG[lx * P2 + k];
                }
                sm.x1[k] = x1;
                sm.x2[k] = x2;
                sm.x12[k] = x1 + x2;
                sm.qcol[k] = qn;
            } else if (warp == 5) {
                // cells of the next column the candidate walk visits: qb != 0, an enclosing pair exists
                const int lx = l - 1;
                int n = 0;
                for (int b = 0; b < PP / 32; b++) {
                    const int k = b * 32 + lane;
                    const bool on = lx > TURN && k >= 1 && k <= lx - TURN - 1 && qbG[lx * P2 + k] != 0.;
                    const unsigned m = __ballot_sync(full, on);
                    if (on) sm.list[lx & 1][n + __popc(m & ((1u << lane) - 1))] = (unsigned char)k;
                    n += __popc(m);
                }
                if (lane == 0) sm.cnt[lx & 1] = n;
            }
            PF2_SYNC(6)
        }

        // ================= ED, centroid =================
        ed_local = warp_sum(ed_local);
        if (lane == 0) sm.red[warp] = ed_local;
        __syncthreads();
        if (tid == 0) {
            double s = 0.;
            for (int w = 0; w < NW2; w++) s += sm.red[w];
            L.ed[fold] = 2. * s;
            L.dG[fold] = (-log(Z) - W * log(T->pf_scale)) * T->kT / 1000.;
        }
        for (int k = tid; k < W; k += NT2) L.centroid[(long long)fold * W + k] = sm.cen[k];
#ifdef SFB_PF2_TIMING
        if (blockIdx.x == 0 && fold == 0 && lane == 0)
            printf("pf2 timing warp %2d: inside p1 %lld wait %lld p2 %lld wait %lld | outside p1 %lld wait %lld p2 %lld wait %lld\n", warp,
                   tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7]);
        if (blockIdx.x == 0 && fold == 0 && lane == 0)
            printf("pf2 probes warp %2d: walk %lld reduce %lld finalize %lld loop %lld\n", warp, tprobe[0], tprobe[1], tprobe[2], tprobe[3]);
#endif
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
    static double kk[32 * 32];
    double scale[40];
    scale[0] = 1.;
    for (int k = 1; k < 40; k++) scale[k] = scale[k - 1] / q.pf_scale;
    for (int u2 = 0; u2 < 32; u2++)
        for (int u1 = 0; u1 < 32; u1++) {
            const int u = u1 + u2, us = u1 < u2 ? u1 : u2, ul = u1 < u2 ? u2 : u1;
            double v = 0.;
            if (u <= MAXLOOP) {
                if (us == 0) {
                    if (ul >= 2) v = q.expbulge[ul] * scale[u + 2];                                 // bulge
                } else if (us == 1) {
                    if (ul >= 3) v = q.expinternal[u] * q.expninio[ul - 1] * scale[u + 2];           // 1xn
                } else if (!(us == 2 && ul <= 3)) {
                    v = q.expinternal[u] * q.expninio[ul - us] * scale[u + 2];                       // generic
                }
            }
            kk[u2 * 32 + u1] = v;
        }
    cudaMemcpyToSymbol(g_K2, kk, sizeof(kk));
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
