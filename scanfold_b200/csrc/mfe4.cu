// MFE fold kernel, fourth generation: BLOCKED int32 fill for long windows (above 300 nt) and whole sequences.
//
// Replaces RNA.fold_compound(seq, md).mfe() where the shared-memory kernels (mfe3.cu) no longer fit:
//   windows above 300 nt of the scan loop      ScanFold.py:494-497 / :512-513 / :534-541, ScanFoldFunctions.py:774-789
//   the full-length refold of --global_refold   ScanFold.py:1518-1539 (hc_add_from_db with the -1 / -2 pairs)
//
// The C / FML / split-minimum matrices of every fold of a batch live in HBM as row-major int32 squares and are
// computed in 32 x 32 blocks.  Block (I, J) only depends on blocks nearer the diagonal, so ONE LAUNCH handles one
// block diagonal (J - I = delta) of every fold of the batch: thousands of independent block tasks, one 256-thread CTA
// each, no barrier between CTAs and nothing to tune per window length.  A task
//   1. accumulates the multiloop splits whose split point lies in the blocks strictly between I and J as min-plus
//      products of 32 x 32 tiles (4 x 4 register tiles, operands staged in shared memory: 2 LDS.128 per 16 VIADDMNMX);
//   2. loads the 62 x 62 window of finished C cells that interior loops of size <= 30 can reach and derives the three
//      copies that carry the inner pair's mismatch term (generic / 1xn / bulge), as mfe.cu's rolling rows do;
//   3. walks the 63 anti-diagonals of the block: per pairable cell one warp spreads the <= 496 interior-loop candidates
//      over its lanes (one LDS + one add-min each), then every cell adds the splits inside the two diagonal blocks.
// HBM traffic is a few tens of KB per task (the matrices are touched once per dependent block, not once per cell), so
// the batch is bound by instruction issue inside the SMs, like the short-window kernels.  Energies are int32 dcal with
// INF = 10^7 exactly as ViennaRNA: no range limit, no redo path.
// A second kernel (one CTA per fold) runs the exterior loop F5 over the transposed copy of C kept in the lower triangle
// and, for native folds, the warp-parallel traceback in ViennaRNA's candidate order (same order as mfe.cu / mfe3.cu).
#include <cstdio>
#include <cstdlib>

#include "device_common.cuh"

namespace sfb {
namespace {

constexpr int BS = 32;        // block edge
constexpr int NT4 = 256;      // threads per block task
constexpr int WR = 62;        // rows / columns of the inner-pair window
constexpr int WP = 64;        // its pitch
constexpr int NCAND = 496;
constexpr int PADG = 416, PAD1 = 64, PADB = 64;   // 406 generic, 54 1xn, 58 bulge candidates, padded to warp multiples

enum { K_GENERIC = 0, K_1N = 1, K_BULGE = 2, K_TABLE = 3 };

struct Tab4 {
    int stack[64];
    int mmI[200], mm1n[200], mm23[200], mmM[200], mmExt[200], mmH[200];
    int d5[40], d3[40];
    int bulge[31], il[31];
    int MLbase, MLclosing, MLintern, ninio, max_ninio, TerminalAU;
    int ncand_upto[32];
    int cand_size[NCAND];
    unsigned short cand_code[NCAND];   // u1 | u2 << 5 | class << 10, sorted by u1 + u2
};

Tab4 *g_dtab4 = nullptr;

struct Mfe4Launch {
    const uint8_t *seqs;   // [n_fold][n] codes (this sub-batch)
    const uint8_t *hc;     // NULL or [n_fold][n] constraint characters
    const int32_t *sc;     // NULL or [n_fold][n+1] 1-based stacking pseudo-energies
    const int32_t *enf;    // NULL or [n_fold][3][n+2]: mate, bal, nxt of the enforced pairs (see enforced_kernel)
    const int32_t *hp_len; // [n+1] hairpin initiation by loop size
    int n_fold, n, NP, NB, max_span;
    int32_t *C, *M, *D;    // [n_fold][NP*NP]: C (upper: C[i][j], lower: its transpose), FML, split minima
    int32_t *e_out;        // [n_fold]
    int16_t *pair16;       // NULL or [n_fold][n]
    int32_t *pair32;       // NULL or [n_fold][n]
    int32_t *tb_stack;     // NULL or [n_fold][3*(n+8)]
};

__device__ __forceinline__ int mmi(int t, int a, int b) { return (t * 5 + a) * 5 + b; }

struct Perm4 {   // pair permission of one fold: codes + constraint flags + enforced pairs
    const uint8_t *S, *hc;
    const int32_t *mate, *bal, *nxt;
    int n, max_span;
    __device__ __forceinline__ int flags(int k) const {
        const char ch = (char)hc[k];
        return (ch == 'x' ? 1 : 0) | (ch == '<' ? 2 : 0) | (ch == '>' ? 4 : 0);
    }
    __device__ int type(int i, int j) const {
        if (j - i <= TURN || i < 0 || j >= n) return 0;
        const int t = pair_type(S[i], S[j]);
        if (!t) return 0;
        if (max_span > 0 && j - i + 1 > max_span) return 0;
        if (hc) {
            const int fi = flags(i), fj = flags(j);
            if ((fi | fj) & 1) return 0;
            if (fj & 2) return 0;   // '<' : j may only pair downstream
            if (fi & 4) return 0;   // '>' : i may only pair upstream
            if (mate) {
                const int mi = mate[i], mj = mate[j];
                if (mi >= 0 && mi != j) return 0;
                if (mj >= 0 && mj != i) return 0;
                // no enforced pair may cross (i,j): the enforced brackets strictly inside must be balanced
                if (mi != j && (bal[j] != bal[i + 1] || nxt[i + 1] <= j)) return 0;
            }
        }
        return t;
    }
};

// enforced pairs '(' ')' of a constraint line (weak enforcement; unbalanced brackets are ignored): per fold
// mate[k] (partner or -1), bal[k] = #openers - #closers among the matched brackets before k, nxt[k] = first k' > k with
// bal[k'] < bal[k] (n + 1 if none).  One thread per fold; the bracket stack lives in the nxt row until it is computed.
__global__ void enforced_kernel(const uint8_t *hc, int n_fold, int n, int32_t *enf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fold) return;
    const uint8_t *h = hc + (size_t)f * n;
    int32_t *mate = enf + (size_t)f * 3 * (n + 2), *bal = mate + (n + 2), *nxt = bal + (n + 2);
    int sp = 0;
    for (int k = 0; k < n + 2; k++) mate[k] = -1;
    for (int k = 0; k < n; k++) {
        if (h[k] == '(')
            nxt[sp++] = k;
        else if (h[k] == ')' && sp > 0) {
            const int a = nxt[--sp];
            mate[a] = k;
            mate[k] = a;
        }
    }
    int b = 0;
    for (int k = 0; k <= n; k++) {
        bal[k] = b;
        if (k < n && mate[k] >= 0) b += mate[k] > k ? 1 : -1;
    }
    bal[n + 1] = b;
    // next smaller value to the right (monotone stack, reusing mate's unused tail is not possible: separate pass)
    for (int k = n + 1; k >= 0; k--) {
        int q = k + 1;
        while (q <= n + 1 && bal[q] >= bal[k]) q = nxt[q];   // nxt[q] is final for q > k
        nxt[k] = q;
    }
}

__device__ int hairpin4(const Tab4 &tb, const MfeTables *T, const int32_t *hp_len, const uint8_t *S, int i, int j, int type) {
    const int u = j - i - 1;
    const int e = hp_len[u];
    if (u < 3) return e;
    if (u == 4) {
        const int key = loop_key_dev(S, i, 6);
        for (int k = 0; k < T->n_tetra; k++)
            if (T->tetra_key[k] == key) return T->tetra_e[k];
    } else if (u == 6) {
        const int key = loop_key_dev(S, i, 8);
        for (int k = 0; k < T->n_hexa; k++)
            if (T->hexa_key[k] == key) return T->hexa_e[k];
    } else if (u == 3) {
        const int key = loop_key_dev(S, i, 5);
        for (int k = 0; k < T->n_tri; k++)
            if (T->tri_key[k] == key) return T->tri_e[k];
        return e + (type > 2 ? tb.TerminalAU : 0);
    }
    return e + tb.mmH[mmi(type, S[i + 1], S[j - 1])];
}

// full interior-loop energy, all classes (the nine table-driven shapes of the fill, every candidate of the traceback)
__device__ int intloop4(const Tab4 &s, const MfeTables *T, int n1, int n2, int type, int t2, int si1, int sj1, int sp1,
                        int sq1) {
    const int nl = max(n1, n2), ns = min(n1, n2);
    if (nl == 0) return s.stack[type * 8 + t2];
    if (ns == 0) {
        int e = s.bulge[nl];
        if (nl == 1)
            e += s.stack[type * 8 + t2];
        else
            e += (type > 2 ? s.TerminalAU : 0) + (t2 > 2 ? s.TerminalAU : 0);
        return e;
    }
    if (ns == 1) {
        if (nl == 1) return T->int11[type][t2][si1][sj1];
        if (nl == 2) return n1 == 1 ? T->int21[type][t2][si1][sq1][sj1] : T->int21[t2][type][sq1][si1][sp1];
        return s.il[nl + 1] + min(s.max_ninio, (nl - ns) * s.ninio) + s.mm1n[mmi(type, si1, sj1)] +
               s.mm1n[mmi(t2, sq1, sp1)];
    }
    if (ns == 2) {
        if (nl == 2) return T->int22[type][t2][si1][sp1][sq1][sj1];
        if (nl == 3) return s.il[5] + s.ninio + s.mm23[mmi(type, si1, sj1)] + s.mm23[mmi(t2, sq1, sp1)];
    }
    return s.il[nl + ns] + min(s.max_ninio, (nl - ns) * s.ninio) + s.mmI[mmi(type, si1, sj1)] + s.mmI[mmi(t2, sq1, sp1)];
}

__device__ __forceinline__ int mlstem4(const Tab4 &s, int type, int si1, int sj1) {
    int e = 0;
    if (si1 >= 0 && sj1 >= 0)
        e = s.mmM[mmi(type, si1, sj1)];
    else if (si1 >= 0)
        e = s.d5[type * 5 + si1];
    else if (sj1 >= 0)
        e = s.d3[type * 5 + sj1];
    return e + (type > 2 ? s.TerminalAU : 0) + s.MLintern;
}

template <class TB>
__device__ __forceinline__ int mlstem4b(const TB &s, int type, int si1, int sj1) {
    int e = 0;
    if (si1 >= 0 && sj1 >= 0)
        e = s.mmM[mmi(type, si1, sj1)];
    else if (si1 >= 0)
        e = s.d5[type * 5 + si1];
    else if (sj1 >= 0)
        e = s.d3[type * 5 + sj1];
    return e + (type > 2 ? s.TerminalAU : 0) + s.MLintern;
}

__device__ __forceinline__ int extloop4(const Tab4 &s, int type, int si1, int sj1) {
    int e = 0;
    if (si1 >= 0 && sj1 >= 0)
        e = s.mmExt[mmi(type, si1, sj1)];
    else if (si1 >= 0)
        e = s.d5[type * 5 + si1];
    else if (sj1 >= 0)
        e = s.d3[type * 5 + sj1];
    return e + (type > 2 ? s.TerminalAU : 0);
}

constexpr int TAB4_INTS = (int)((sizeof(Tab4) + 15) / 16 * 4);

// What the block kernel keeps in shared memory (the rarely used tables stay in the global Tab4): the interior-loop
// candidates of the three separable classes as ONE word each -- window offset u1 * WP - u2 + 30 in the low half, size term
// minus the field bias in the high half -- sorted by class, then by loop size u1 + u2, so that a class is one loop with a
// compile-time field shift; the nine table-driven shapes run apart.
struct TabB {
    int stack[64];
    int mmI[200], mm1n[200], mmM[200];
    int d5[40], d3[40];
    int mm23[200], mmH[200];
    int MLbase, MLclosing, MLintern, TerminalAU, bulge1, il5_ninio, pad1, pad2;
    int cls_begin[4];          // first candidate of class 0, 1, 2 (and the end)
    int ncls_upto[3][32];      // candidates of the class with u1 + u2 <= u
    // ---- from here on the block kernel reads the global copy (start-up and the short walks of small block diagonals)
    int cand[NCAND];
    int candp[PADG + PAD1 + PADB];   // the same lists, each padded to a multiple of 32 with never-winning entries: the full
                                     // walk (every cell with j - i > 35) is fixed-trip and each lane keeps its 17 words in registers
};
TabB *g_dtabB = nullptr;
bool g_mfe4_ok = false;

constexpr int TABB_INTS = (int)(offsetof(TabB, cand) / 4);   // the part copied to shared memory
constexpr int TRI = 32 * 33 / 2;                  // upper triangle of a 32 x 32 block, diagonal included
constexpr int O_MB = TABB_INTS;                   // [33][34] FML: row 32 = first row of block (I+1,J), column 0 = column 32J-1
constexpr int O_DB = O_MB + 33 * 34;              // [33][34] split minima, same halo
constexpr int O_WIN = O_DB + 33 * 34;             // int2 [WR][WP]: C and the packed mismatch fields of the inner-pair window;
                                                  // from here on the area doubles as the tile staging of step 1
constexpr int O_MII = O_WIN + 2 * WR * WP;        // [TRI] FML of the diagonal block (I,I), upper triangle
constexpr int O_MJJ = O_MII + TRI;                // [TRI] FML of the diagonal block (J,J)
constexpr int O_PART = O_MJJ + TRI;               // [2][2][32] partial minima of the cells of a step: separable classes / shapes
constexpr int O_TYB = O_PART + 128;               // bytes [32][32]: pair type of every cell of the block (0 = may not pair)
constexpr int O_SLIST = O_TYB + 256;              // bytes [63][32]: pairable cells (column b) of every step, compacted
constexpr int O_SCNT = O_SLIST + 63 * 8;          // bytes [64]: their number
constexpr int O_SEQ = O_SCNT + 16;                // bytes [2][160]: row / column slices of the sequence
constexpr int O_END = O_SEQ + 80;
constexpr size_t SMEM4_BYTES = (size_t)O_END * 4;
static_assert(4 * (32 * 36 + 32 * 32) <= O_END - O_WIN, "the staging tiles of step 1 fit behind the accumulators");
static_assert(4 * (SMEM4_BYTES + 1024) <= 227 * 1024, "four block tasks per SM");

__device__ __forceinline__ int tri32(int r, int c) { return r * 32 - (r * (r - 1)) / 2 + (c - r); }   // c >= r

constexpr int FB = 512;                           // bias of a packed 10-bit mismatch field
__device__ __forceinline__ int pack3(int a, int b, int c) { return (a + FB) | ((b + FB) << 10) | ((c + FB) << 20); }

__global__ void __launch_bounds__(NT4, 4)
mfe4_block_kernel(Mfe4Launch L, const MfeTables *__restrict__ T, const Tab4 *__restrict__ gtab, const TabB *__restrict__ gtabB,
                  int delta) {
    extern __shared__ __align__(16) int sm4[];
    TabB &tb = *reinterpret_cast<TabB *>(sm4);
    int *Mb = sm4 + O_MB, *Db = sm4 + O_DB, *Mii = sm4 + O_MII, *Mjj = sm4 + O_MJJ, *part = sm4 + O_PART;
    int2 *win = reinterpret_cast<int2 *>(sm4 + O_WIN);
    unsigned char *tyb = reinterpret_cast<unsigned char *>(sm4 + O_TYB);
    unsigned char *slist = reinterpret_cast<unsigned char *>(sm4 + O_SLIST);
    unsigned char *scnt = reinterpret_cast<unsigned char *>(sm4 + O_SCNT);
    unsigned char *sR = reinterpret_cast<unsigned char *>(sm4 + O_SEQ);   // codes of positions 32I-1 .. 32I+158
    unsigned char *sC = sR + 160;                                        // codes of positions 32J-33 .. 32J+126
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned full = 0xffffffffu;
    {
        const int *src = reinterpret_cast<const int *>(gtabB);
        for (int k = tid; k < TABB_INTS; k += NT4) sm4[k] = src[k];
    }
    __syncthreads();
    int cdr[(PADG + PAD1 + PADB) / 32];   // this lane's candidate words of the full walk (warps 2..6 use them)
#pragma unroll
    for (int k = 0; k < (PADG + PAD1 + PADB) / 32; k++) cdr[k] = gtabB->candp[k * 32 + lane];
    const int n = L.n, NP = L.NP;
    const int per_fold = L.NB - delta;   // block tasks of this diagonal per fold
    const long long n_task = (long long)L.n_fold * per_fold;
    for (long long task = blockIdx.x; task < n_task; task += gridDim.x) {
        const int fold = (int)(task / per_fold), I = (int)(task % per_fold), J = I + delta;
        const int i0 = I * BS, j0 = J * BS;
        const uint8_t *S = L.seqs + (size_t)fold * n;
        const size_t moff = (size_t)fold * NP * NP;
        int32_t *gC = L.C + moff, *gM = L.M + moff, *gD = L.D + moff;
#ifdef SFB_MFE4_TIMING
        const long long tm_task0 = clock64();
#endif
        Perm4 pm;
        pm.S = S;
        pm.hc = L.hc ? L.hc + (size_t)fold * n : nullptr;
        pm.mate = L.enf ? L.enf + (size_t)fold * 3 * (n + 2) : nullptr;
        pm.bal = pm.mate ? pm.mate + (n + 2) : nullptr;
        pm.nxt = pm.mate ? pm.bal + (n + 2) : nullptr;
        pm.n = n;
        pm.max_span = L.max_span;
        const int32_t *scf = L.sc ? L.sc + (size_t)fold * (n + 1) : nullptr;
        __syncthreads();   // previous task done with shared memory
        // guarded matrix loads: cells outside the triangle (or the fold) read as INF
        auto ldM = [&](const int32_t *g, int r, int c) { return (r >= 0 && c < n && c - r > TURN) ? __ldcg(g + (size_t)r * NP + c) : INF; };
        for (int k = tid; k < 33 * 34; k += NT4) {
            Db[k] = INF;
            Mb[k] = INF;
        }
        for (int k = tid; k < 160; k += NT4) {
            const int pr = i0 - 1 + k, pc = j0 - 33 + k;
            sR[k] = (pr >= 0 && pr < n) ? S[pr] : 0;
            sC[k] = (pc >= 0 && pc < n) ? S[pc] : 0;
        }
        __syncthreads();
        auto SR = [&](int p) -> int { return sR[p - (i0 - 1)]; };
        auto SCc = [&](int q) -> int { return sC[q - (j0 - 33)]; };

        // ---- 1. splits with the split point in blocks strictly between I and J: min-plus tile products
        if (delta >= 2) {
            const int wp = warp >> 1, half = warp & 1;
            const int rb = 16 * half + 4 * (lane >> 3), cb = 4 * (lane & 7);
            int *At = sm4 + O_WIN + wp * (32 * 36 + 32 * 32), *Bs = At + 32 * 36;
            int acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = 2 * INF;
            const int tp = half * 32 + lane;
            for (int K0 = I + 1; K0 < J; K0 += 4) {
                const int K = K0 + wp;
                __syncthreads();
                if (K < J) {
                    const int k0 = K * BS;
#pragma unroll 4
                    for (int r = 0; r < 16; r++) {
                        const int a = 2 * r + (tp >> 5), c = tp & 31;
                        // left operand L[i][m] = FML[i][m-1], m = k0 + c; right operand FML[m][j]
                        At[c * 36 + a] = ldM(gM, i0 + a, k0 + c - 1);
                        Bs[a * 32 + c] = ldM(gM, k0 + a, j0 + c);
                    }
                }
                __syncthreads();
                if (K < J) {
#pragma unroll 8
                    for (int m = 0; m < 32; m++) {
                        const int4 av = *reinterpret_cast<const int4 *>(At + m * 36 + rb);
                        const int4 bv = *reinterpret_cast<const int4 *>(Bs + m * 32 + cb);
                        const int aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int b = 0; b < 4; b++) acc[a][b] = __viaddmin_s32(aa[a], bb[b], acc[a][b]);
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (acc[a][b] < INF / 2) atomicMin(&Db[(rb + a) * 34 + cb + b + 1], acc[a][b]);
            __syncthreads();
        }

        // ---- 2. halos, diagonal blocks, pair types and step lists, inner-pair window
        for (int k = tid; k < 34; k += NT4) {   // row 32 (first row of block (I+1, J)), columns 32J-1 .. 32J+32
            Mb[32 * 34 + k] = ldM(gM, i0 + 32, j0 - 1 + k);
            Db[32 * 34 + k] = ldM(gD, i0 + 32, j0 - 1 + k);
        }
        for (int k = tid; k < 32; k += NT4) {   // column 0 (column 32J-1 of block (I, J-1))
            Mb[k * 34] = ldM(gM, i0 + k, j0 - 1);
            Db[k * 34] = ldM(gD, i0 + k, j0 - 1);
        }
        if (delta >= 1)
            for (int k = tid; k < 32 * 32; k += NT4) {
                const int a = k >> 5, c = k & 31;
                if (c >= a) {   // Mii by row (entry [a][c] at tri32(a, c)), Mjj by column (entry [a][c] at c (c + 1) / 2 + a)
                    Mii[tri32(a, c)] = ldM(gM, i0 + a, i0 + c);
                    Mjj[(c * (c + 1)) / 2 + a] = ldM(gM, j0 + a, j0 + c);
                }
            }
        for (int k = tid; k < 32 * 32; k += NT4) {
            const int a = k >> 5, b = k & 31;
            tyb[k] = (unsigned char)pm.type(i0 + a, j0 + b);
        }
        for (int k = tid; k < 128; k += NT4) part[k] = INF;
        {
            const int r0 = i0 + 1, c0 = j0 - 31;
            const int none = pack3(0, 0, 0);
            for (int k = tid; k < WR * WR; k += NT4) {
                const int wr = k / WR, wc = k - wr * WR;
                const int p = r0 + wr, q = c0 + wc;
                int2 e = make_int2(INF, none);
                const bool own = wr <= 30 && wc >= 31;   // cells of this block: filled in step 3
                if (!own && q >= 1 && q < n - 1 && p < n && q - p > TURN) {
                    const int c = __ldcg(gC + (size_t)p * NP + q);
                    if (c < INF) {
                        const int t2 = rtype_of(pair_type(SR(p), SCc(q)));
                        const int m2 = mmi(t2, SCc(q + 1), SR(p - 1));
                        e = make_int2(c, pack3(tb.mmI[m2], tb.mm1n[m2], t2 > 2 ? tb.TerminalAU : 0));
                    }
                }
                win[wr * WP + wc] = e;
            }
        }
        __syncthreads();
        const int s_begin = delta == 0 ? 31 + TURN + 1 : 0;   // diagonal blocks: j - i <= TURN before that
        for (int s = s_begin + warp; s < 63; s += NT4 / 32) {   // compacted list of the pairable cells of every step
            const int blo = max(0, s - 31), nc = min(s, 62 - s) + 1;
            const int b = blo + lane, a = 31 - s + b;
            const int t = lane < nc ? (int)tyb[a * 32 + b] : 0;
            const unsigned m = __ballot_sync(full, t != 0);
            if (t) slist[s * 32 + __popc(m & ((1u << lane) - 1))] = (unsigned char)b;
            if (lane == 0) scnt[s] = (unsigned char)__popc(m);
        }
        __syncthreads();

        // ---- 3. the 63 anti-diagonals of the block, cell (a, b) = (i0 + a, j0 + b) on step s = 31 - a + b, software
        // pipelined with ONE barrier per step.  Iteration t runs, side by side,
        //   G  the separable interior loops of the pairable cells of step t, one warp per cell (warps 2..6)
        //   H  their nine table-driven shapes and the hairpin, lane = cell (warps 0 and 1 share the shapes)
        //   F  for the cells of step t-1: C from the partial minima of iteration t-1 plus the multiloop closing term and the
        //      window entry (warp 7, lane = cell), the splits inside the two diagonal blocks (warps 5 and 6, before their
        //      cells), then FML once the three warps met at a named barrier (warp 7)
        // C(t) only reads cells of steps <= t-2 and the split minima of step t-2; FML(t-1) reads FML of steps <= t-2.
#ifdef SFB_MFE4_TIMING
        long long tm_busy = 0, tm_loop0 = clock64(), tm_a = 0, tm_b = 0;
#endif
        for (int t = s_begin; t <= 63; t++) {
#ifdef SFB_MFE4_TIMING
            const long long tm_it0 = clock64();
#endif
            if (warp == NT4 / 32 - 1) {
                // ---- F: step t - 1, lane = cell
                if (t > s_begin) {
                    const int s = t - 1;
                    const int blo = max(0, s - 31), nc = min(s, 62 - s) + 1;
                    const int b = blo + min(lane, nc - 1), a = 31 - s + b;
                    const int i = i0 + a, j = j0 + b;
                    const bool active = lane < nc && j < n;
                    int cij = INF;
                    if (active) {
                        const int type = tyb[a * 32 + b];
                        if (type) {
                            int e = min(part[(s & 1) * 64 + b], part[(s & 1) * 64 + 32 + b]);
                            part[(s & 1) * 64 + b] = INF;        // both are written again two iterations on
                            part[(s & 1) * 64 + 32 + b] = INF;
                            if (j - i >= 2 + TURN + 1 + 2) {   // multiloop closed by (i,j): split minimum of cell (i+1, j-1)
                                const int dm = Db[(a + 1) * 34 + b];
                                if (dm < INF)
                                    e = min(e, dm + mlstem4b(tb, rtype_of(type), SCc(j - 1), SR(i + 1)) + tb.MLclosing);
                            }
                            cij = e;
                            if (a >= 1 && b <= 30 && i > 0 && j < n - 1) {   // inner pair of later cells of this block
                                const int t2 = rtype_of(type);
                                const int m2 = mmi(t2, SCc(j + 1), SR(i - 1));
                                win[(a - 1) * WP + b + 31] = make_int2(e, pack3(tb.mmI[m2], tb.mm1n[m2], t2 > 2 ? tb.TerminalAU : 0));
                            }
                        }
                        if (j - i > TURN) {
                            gC[(size_t)i * NP + j] = cij;
                            gC[(size_t)j * NP + i] = cij;   // transposed copy for the exterior loop
                        }
                    }
                    // the splits inside the two diagonal blocks come from warps 5 and 6 (below) through Db
                    asm volatile("bar.sync 1, 96;" ::: "memory");
                    int dec = 2 * INF;
                    if (active) {
                        dec = min(dec, Db[a * 34 + b + 1]);
                        if (dec > INF / 2) dec = INF;
                        int m = dec;
                        if (j - i - 1 > TURN) {
                            const int x = Mb[(a + 1) * 34 + b + 1], y = Mb[a * 34 + b];
                            if (x < INF) m = min(m, x + tb.MLbase);
                            if (y < INF) m = min(m, y + tb.MLbase);
                        }
                        if (cij < INF)
                            m = min(m, cij + mlstem4b(tb, pair_type(SR(i), SCc(j)), i > 0 ? SR(i - 1) : -1, j < n - 1 ? SCc(j + 1) : -1));
                        if (j - i <= TURN) m = INF;
                        Db[a * 34 + b + 1] = dec;
                        Mb[a * 34 + b + 1] = m;
                        if (j - i > TURN) {
                            gM[(size_t)i * NP + j] = m;
                            gD[(size_t)i * NP + j] = dec;
                        }
                    }
                }
            } else {
              if (warp >= 5 && t > s_begin) {   // the two interior-loop warps with the fewest cells (round robin from warp 2)
                // ---- F, split part: the splits of step t - 1 inside the two diagonal blocks, lane = cell.  m = i0 + c (rows
                // below, 31 - a terms: warp 5) and m = j0 + c (columns to the left, b + 1 terms: warp 6).  Both loops run over
                // warp-uniform ranges WITHOUT a per-lane test: outside a lane's own range one operand is a cell of this block
                // that a later step computes -- still INF in Mb -- so the sum stays above INF / 2 whatever the other operand
                // reads (neighbouring entries of the triangles).  The minima meet the far splits in Db (shared-memory
                // atomics, one cell per lane); warp 7 picks them up behind the named barrier.
                const int s = t - 1;
                const int blo = max(0, s - 31), nc = min(s, 62 - s) + 1;
                const int b = blo + min(lane, nc - 1), a = 31 - s + b;
                const int a_lo = 31 - s + blo, b_hi = blo + nc - 1;   // smallest a / largest b of the step
                int dec = 2 * INF;
                if (delta == 0) {   // a < c <= b: FML[i][m-1] + FML[m][j], both inside this block
                    if (warp == 5) {
                        const int *pa = Mb + a * 34, *pb = Mb + b + 1;
#pragma unroll 4
                        for (int c = a_lo + 1; c <= b_hi; c++) dec = __viaddmin_s32(pa[c], pb[c * 34], dec);
                    }
                } else if (warp == 5) {
                    const int *pa = Mii + tri32(a, a) - a - 1, *pb = Mb + b + 1;   // pa[c] = Mii[a][c-1]
#pragma unroll 4
                    for (int c = a_lo + 1; c < 32; c++) dec = __viaddmin_s32(pa[c], pb[c * 34], dec);
                } else {
                    const int *qa = Mb + a * 34, *qb = Mjj + (b * (b + 1)) / 2;   // Mjj by column: [b][c], c <= b
#pragma unroll 4
                    for (int c = 0; c <= b_hi; c++) dec = __viaddmin_s32(qa[c], qb[c], dec);
                }
                if (lane < nc && dec < INF / 2) atomicMin(&Db[a * 34 + b + 1], dec);
#ifdef SFB_MFE4_TIMING
                tm_a += clock64() - tm_it0;
                const long long tm_b0 = clock64();
#endif
                asm volatile("bar.sync 1, 96;" ::: "memory");
#ifdef SFB_MFE4_TIMING
                tm_b += clock64() - tm_b0;
#endif
              }
              if (t < 63) {
                const int npair = scnt[t];
                const unsigned char *lst = slist + t * 32;
                int *pG = part + (t & 1) * 64, *pS = pG + 32;
                if (warp >= 2) {
                    // ---- G: separable classes, one warp per cell, lane = candidate, one loop per class
                    for (int c = warp - 2; c < npair; c += NT4 / 32 - 3) {
                        const int b = lst[c], a = 31 - t + b;
                        const int type = tyb[a * 32 + b];
                        const int i = i0 + a, j = j0 + b, d = j - i;
                        const int mi = mmi(type, SR(i + 1), SCc(j - 1));
                        const int umax = min(MAXLOOP, d - 2 - (TURN + 1));
                        const int2 *base = win + a * WP + b;
                        int accG = INF, acc1 = INF, accB = INF;
                        if (umax == MAXLOOP) {   // the full candidate set: fixed trip counts, candidate words from registers
#pragma unroll
                            for (int k = 0; k < PADG / 32; k++) {
                                const int2 w = base[cdr[k] & 0xffff];
                                accG = __viaddmin_s32(w.x + (w.y & 1023), cdr[k] >> 16, accG);
                            }
#pragma unroll
                            for (int k = 0; k < PAD1 / 32; k++) {
                                const int cd = cdr[PADG / 32 + k];
                                const int2 w = base[cd & 0xffff];
                                acc1 = __viaddmin_s32(w.x + ((w.y >> 10) & 1023), cd >> 16, acc1);
                            }
#pragma unroll
                            for (int k = 0; k < PADB / 32; k++) {
                                const int cd = cdr[(PADG + PAD1) / 32 + k];
                                const int2 w = base[cd & 0xffff];
                                accB = __viaddmin_s32(w.x + ((w.y >> 20) & 1023), cd >> 16, accB);
                            }
                        } else if (umax >= 0) {
                            const int *c0 = gtabB->cand + tb.cls_begin[0], *c1 = gtabB->cand + tb.cls_begin[1], *c2 = gtabB->cand + tb.cls_begin[2];
                            const int n0 = tb.ncls_upto[0][umax], n1 = tb.ncls_upto[1][umax], n2 = tb.ncls_upto[2][umax];
                            for (int ci = lane; ci < n0; ci += 32) {
                                const int cd = c0[ci];
                                const int2 w = base[cd & 0xffff];
                                accG = __viaddmin_s32(w.x + (w.y & 1023), cd >> 16, accG);
                            }
                            for (int ci = lane; ci < n1; ci += 32) {
                                const int cd = c1[ci];
                                const int2 w = base[cd & 0xffff];
                                acc1 = __viaddmin_s32(w.x + ((w.y >> 10) & 1023), cd >> 16, acc1);
                            }
                            for (int ci = lane; ci < n2; ci += 32) {
                                const int cd = c2[ci];
                                const int2 w = base[cd & 0xffff];
                                accB = __viaddmin_s32(w.x + ((w.y >> 20) & 1023), cd >> 16, accB);
                            }
                        }
                        int acc = min(accG + tb.mmI[mi], min(acc1 + tb.mm1n[mi], accB + (type > 2 ? tb.TerminalAU : 0)));
                        acc = __reduce_min_sync(full, acc);
                        if (lane == 0) pG[b] = min(acc, INF);
                    }
                } else if (lane < npair) {
                    // ---- H: the nine table-driven shapes (warp 0: stack, bulges of one, 1x1, hairpin; warp 1: 1x2, 2x1, 2x2,
                    // 2x3, 3x2), lane = cell.  Straight-line code: every table load is issued whether or not the inner cell
                    // pairs (type 0 rows are valid addresses), so the L2 round trips of a warp overlap.
                    const int b = lst[lane], a = 31 - t + b;
                    const int type = tyb[a * 32 + b];
                    const int i = i0 + a, j = j0 + b;
                    const int si1 = SR(i + 1), sj1 = SCc(j - 1);
                    const int2 *wb = win + a * WP + b + 30;
                    auto inner = [&](int u1, int u2, int &cpq, int &t2, int &sp1, int &sq1) {
                        cpq = wb[u1 * WP - u2].x;
                        const int p = i + 1 + u1, q = j - 1 - u2;
                        t2 = rtype_of(pair_type(SR(p), SCc(q)));
                        sp1 = SR(p - 1);
                        sq1 = SCc(q + 1);
                    };
                    auto fin = [](int c, int e) { return c < INF ? c + e : INF; };
                    int best;
                    if (warp == 0) {
                        int c00, c01, c10, c11, t00, t01, t10, t11, x, y;
                        inner(0, 0, c00, t00, x, y);
                        inner(0, 1, c01, t01, x, y);
                        inner(1, 0, c10, t10, x, y);
                        inner(1, 1, c11, t11, x, y);
                        const int e11 = __ldg(&T->int11[type][t11][si1][sj1]);
                        const int u = j - i - 1;
                        int eh = __ldg(L.hp_len + u) + tb.mmH[mmi(type, si1, sj1)];
                        if (u <= 6) eh = hairpin4(*gtab, T, L.hp_len, S, i, j, type);
                        int e00 = tb.stack[type * 8 + t00];
                        if (scf) e00 += scf[i + 1] + scf[i + 2] + scf[j] + scf[j + 1];
                        best = min(eh, fin(c00, e00));
                        best = min(best, fin(c01, tb.bulge1 + tb.stack[type * 8 + t01]));
                        best = min(best, fin(c10, tb.bulge1 + tb.stack[type * 8 + t10]));
                        best = min(best, fin(c11, e11));
                    } else {
                        int c12, c21, c22, c23, c32, t12, t21, t22, t23, t32, p12, q12, p21, q21, p22, q22, p23, q23, p32, q32;
                        inner(1, 2, c12, t12, p12, q12);
                        inner(2, 1, c21, t21, p21, q21);
                        inner(2, 2, c22, t22, p22, q22);
                        inner(2, 3, c23, t23, p23, q23);
                        inner(3, 2, c32, t32, p32, q32);
                        const int e12 = __ldg(&T->int21[type][t12][si1][q12][sj1]);     // n1 == 1
                        const int e21 = __ldg(&T->int21[t21][type][q21][si1][p21]);
                        const int e22 = __ldg(&T->int22[type][t22][si1][p22][q22][sj1]);
                        const int m23 = tb.il5_ninio + tb.mm23[mmi(type, si1, sj1)];
                        best = min(fin(c12, e12), fin(c21, e21));
                        best = min(best, fin(c22, e22));
                        best = min(best, fin(c23, m23 + tb.mm23[mmi(t23, q23, p23)]));
                        best = min(best, fin(c32, m23 + tb.mm23[mmi(t32, q32, p32)]));
                    }
                    atomicMin(&pS[b], best);
                }
              }
            }
#ifdef SFB_MFE4_TIMING
            tm_busy += clock64() - tm_it0;
#endif
            __syncthreads();
        }
#ifdef SFB_MFE4_TIMING
        if (delta == 8 && blockIdx.x < 3 && task == blockIdx.x && lane == 0)
            printf("mfe4 timing cta %d warp %d: busy %lld of loop %lld cycles (%d iterations), phases 1-2 %lld, splits %lld, named barrier %lld\n",
                   blockIdx.x, warp, tm_busy, clock64() - tm_loop0, 64 - s_begin, tm_loop0 - tm_task0, tm_a, tm_b);
#endif
    }
}

// ---- exterior loop and traceback: one CTA per fold
constexpr int NT4B = 128;

__device__ bool traceback4(const Mfe4Launch &L, const Tab4 &tb, const MfeTables *T, const uint8_t *S, const int32_t *scf,
                           const int32_t *gC, const int32_t *gM, const int *f5, int *stk, int fold, int lane) {
    const int n = L.n, NP = L.NP;
    const unsigned full = 0xffffffffu;
    auto CC = [&](int i, int j) { return (j - i > TURN) ? __ldcg(gC + (size_t)i * NP + j) : INF; };
    auto MM = [&](int i, int j) { return (j - i > TURN) ? __ldcg(gM + (size_t)i * NP + j) : INF; };
    auto nb5 = [&](int i) { return i > 0 ? (int)S[i - 1] : -1; };
    auto nb3 = [&](int j) { return j < n - 1 ? (int)S[j + 1] : -1; };
    auto setpair = [&](int i, int j) {
        if (L.pair16) {
            L.pair16[(size_t)fold * n + i] = (int16_t)(j + 1);
            L.pair16[(size_t)fold * n + j] = (int16_t)(i + 1);
        } else {
            L.pair32[(size_t)fold * n + i] = j + 1;
            L.pair32[(size_t)fold * n + j] = i + 1;
        }
    };
    int sp = 1;
    if (lane == 0) {
        stk[0] = 0;
        stk[1] = n - 1;
        stk[2] = 0;
    }
    __syncwarp();
    while (sp > 0) {
        sp--;
        int i = stk[3 * sp], j = stk[3 * sp + 1];
        const int ml = stk[3 * sp + 2];
        __syncwarp();
        bool have_pair = false;
        if (j < i + TURN + 1) continue;
        const int fij = ml ? MM(i, j) : f5[j + 1];
        const int mij1 = MM(i, j - 1);
        const int fi = ml ? (mij1 < INF ? mij1 + tb.MLbase : INF) : f5[j];
        auto push = [&](int a, int b, int c) {
            if (lane == 0) {
                stk[3 * sp] = a;
                stk[3 * sp + 1] = b;
                stk[3 * sp + 2] = c;
            }
            sp++;
        };
        if (fij == fi) {
            push(i, j - 1, ml);
            __syncwarp();
            continue;
        }
        if (ml == 0) {
            int kf = -1;
            for (int k0 = j - TURN - 1; k0 >= 0 && kf < 0; k0 -= 32) {
                const int k = k0 - lane;
                bool hit = false;
                if (k >= 0) {
                    const int ckj = CC(k, j);
                    hit = ckj < INF && fij == extloop4(tb, pair_type(S[k], S[j]), nb5(k), nb3(j)) + ckj + f5[k];
                }
                const unsigned m = __ballot_sync(full, hit);
                if (m) kf = k0 - (__ffs(m) - 1);
            }
            if (kf < 0) return false;
            push(0, kf - 1, 0);
            i = kf;
            have_pair = true;
        } else {
            const int mi1j = MM(i + 1, j);
            if (mi1j < INF && mi1j + tb.MLbase == fij) {
                push(i + 1, j, 1);
                __syncwarp();
                continue;
            }
            const int cij = CC(i, j);
            if (cij < INF && fij == cij + mlstem4(tb, pair_type(S[i], S[j]), nb5(i), nb3(j))) {
                have_pair = true;
            } else {
                int kf = -1;
                for (int k0 = i + 1 + TURN; k0 <= j - 2 - TURN && kf < 0; k0 += 32) {
                    const int k = k0 + lane;
                    bool hit = false;
                    if (k <= j - 2 - TURN) {
                        const int a = MM(i, k), b = MM(k + 1, j);
                        hit = a < INF && b < INF && fij == a + b;
                    }
                    const unsigned m = __ballot_sync(full, hit);
                    if (m) kf = k0 + __ffs(m) - 1;
                }
                if (kf < 0) return false;
                push(i, kf, 1);
                push(kf + 1, j, 1);
                __syncwarp();
                continue;
            }
        }
        while (have_pair) {
            if (lane == 0) setpair(i, j);
            const int type = pair_type(S[i], S[j]);
            const int cij = CC(i, j);
            if (cij == hairpin4(tb, T, L.hp_len, S, i, j, type)) break;
            bool traced = false;
            const int pmax = min(j - 2 - TURN, i + MAXLOOP + 1);
            for (int p = i + 1; p <= pmax && !traced; p++) {
                int minq = j - i + p - MAXLOOP - 2;
                if (minq < p + 1 + TURN) minq = p + 1 + TURN;
                const int q = j - 1 - lane;   // at most 31 candidates per p
                bool hit = false;
                if (q >= minq) {
                    const int cpq = CC(p, q);
                    if (cpq < INF) {
                        const int t2 = rtype_of(pair_type(S[p], S[q]));
                        int e = intloop4(tb, T, p - i - 1, j - q - 1, type, t2, S[i + 1], S[j - 1], S[p - 1], S[q + 1]);
                        if (p == i + 1 && q == j - 1 && scf) e += scf[i + 1] + scf[p + 1] + scf[q + 1] + scf[j + 1];
                        hit = cij == e + cpq;
                    }
                }
                const unsigned m = __ballot_sync(full, hit);
                if (m) {
                    j = j - 1 - (__ffs(m) - 1);
                    i = p;
                    traced = true;
                }
            }
            if (traced) continue;
            const int en = cij - mlstem4(tb, rtype_of(type), S[j - 1], S[i + 1]) - tb.MLclosing;
            int kf = -1;
            for (int k0 = i + 2 + TURN; k0 < j - 2 - TURN && kf < 0; k0 += 32) {
                const int k = k0 + lane;
                bool hit = false;
                if (k < j - 2 - TURN) {
                    const int a = MM(i + 1, k), b = MM(k + 1, j - 1);
                    hit = a < INF && b < INF && en == a + b;
                }
                const unsigned m = __ballot_sync(full, hit);
                if (m) kf = k0 + __ffs(m) - 1;
            }
            if (kf < 0) return false;
            push(i + 1, kf, 1);
            push(kf + 1, j - 1, 1);
            break;
        }
        __syncwarp();
    }
    return true;
}

__global__ void __launch_bounds__(NT4B) mfe4_exterior_kernel(Mfe4Launch L, const MfeTables *__restrict__ T,
                                                              const Tab4 *__restrict__ gtab) {
    extern __shared__ __align__(16) int smx[];
    Tab4 &tb = *reinterpret_cast<Tab4 *>(smx);
    int *red = smx + TAB4_INTS;        // [8]
    int *f5 = red + 8;                 // [n + 2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    {
        const int *src = reinterpret_cast<const int *>(gtab);
        for (int k = tid; k < (int)(sizeof(Tab4) / 4); k += NT4B) smx[k] = src[k];
    }
    const int n = L.n, NP = L.NP;
    for (int fold = blockIdx.x; fold < L.n_fold; fold += gridDim.x) {
        const uint8_t *S = L.seqs + (size_t)fold * n;
        const int32_t *gC = L.C + (size_t)fold * NP * NP, *gM = L.M + (size_t)fold * NP * NP;
        __syncthreads();
        for (int k = tid; k <= min(n, TURN + 1); k += NT4B) f5[k] = 0;
        __syncthreads();
        for (int len = TURN + 2; len <= n; len++) {
            const int j = len - 1;
            const int32_t *row = gC + (size_t)j * NP;   // transposed copy: row[i] = C[i][j]
            const int sj1 = j < n - 1 ? (int)S[j + 1] : -1;
            int best = INF;
            for (int i = tid; i <= j - TURN - 1; i += NT4B) {
                const int cij = __ldcg(row + i);
                if (cij < INF) best = min(best, f5[i] + cij + extloop4(tb, pair_type(S[i], S[j]), i > 0 ? (int)S[i - 1] : -1, sj1));
            }
            best = __reduce_min_sync(full, best);
            if (lane == 0) red[warp] = best;
            __syncthreads();
            if (tid == 0) {
                int b = f5[len - 1];
                for (int w = 0; w < NT4B / 32; w++) b = min(b, red[w]);
                f5[len] = b;
            }
            __syncthreads();
        }
        if (tid == 0) L.e_out[fold] = f5[n];
        if (L.pair16 || L.pair32) {
            for (int k = tid; k < n; k += NT4B) {
                if (L.pair16) L.pair16[(size_t)fold * n + k] = 0;
                else L.pair32[(size_t)fold * n + k] = 0;
            }
            __syncthreads();
            if (warp == 0) {
                const int32_t *scf = L.sc ? L.sc + (size_t)fold * (n + 1) : nullptr;
                const bool ok = traceback4(L, tb, T, S, scf, gC, gM, f5, L.tb_stack + (size_t)fold * 3 * (n + 8), fold, lane);
                if (!ok && lane == 0) L.e_out[fold] = INF;   // surfaces as an error on the host
            }
        }
    }
}

}  // namespace

void mfe4_upload_tables(const MfeTables &M) {
    static Tab4 h;
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++) h.stack[a * 8 + b] = M.stack[a][b];
    for (int k = 0; k < 200; k++) {
        h.mmI[k] = (&M.mismatchI[0][0][0])[k];
        h.mm1n[k] = (&M.mismatch1nI[0][0][0])[k];
        h.mm23[k] = (&M.mismatch23I[0][0][0])[k];
        h.mmM[k] = (&M.mismatchM[0][0][0])[k];
        h.mmExt[k] = (&M.mismatchExt[0][0][0])[k];
        h.mmH[k] = (&M.mismatchH[0][0][0])[k];
    }
    for (int k = 0; k < 40; k++) {
        h.d5[k] = (&M.dangle5[0][0])[k];
        h.d3[k] = (&M.dangle3[0][0])[k];
    }
    for (int k = 0; k < 31; k++) {
        h.bulge[k] = M.bulge[k];
        h.il[k] = M.internal_loop[k];
    }
    h.MLbase = M.MLbase;
    h.MLclosing = M.MLclosing;
    h.MLintern = M.MLintern;
    h.ninio = M.ninio;
    h.max_ninio = M.max_ninio;
    h.TerminalAU = M.TerminalAU;
    int nc = 0;
    for (int u = 0; u <= MAXLOOP; u++) {
        for (int u1 = 0; u1 <= u; u1++) {
            const int u2 = u - u1, nl = u1 > u2 ? u1 : u2, ns = u1 > u2 ? u2 : u1;
            int cls = K_TABLE, size = 0;
            if (ns == 0 && nl >= 2) {
                cls = K_BULGE;
                size = M.bulge[nl];
            } else if (ns == 1 && nl >= 3) {
                cls = K_1N;
                size = M.internal_loop[nl + 1] + std::min(M.max_ninio, (nl - ns) * M.ninio);
            } else if (ns >= 2 && !(ns == 2 && nl <= 3)) {
                cls = K_GENERIC;
                size = M.internal_loop[u] + std::min(M.max_ninio, (nl - ns) * M.ninio);
            }
            h.cand_code[nc] = (unsigned short)(u1 | (u2 << 5) | (cls << 10));
            h.cand_size[nc] = size;
            nc++;
        }
        h.ncand_upto[u] = nc;
    }
    h.ncand_upto[31] = nc;
    if (!g_dtab4) cudaMalloc(&g_dtab4, sizeof(Tab4));
    cudaMemcpy(g_dtab4, &h, sizeof(Tab4), cudaMemcpyHostToDevice);

    // the block kernel's shared-memory copy: packed candidates of the separable classes only
    static TabB hb;
    bool ok = true;
    auto fits10 = [&](int v) { return v >= -FB && v < FB; };
    for (int k = 0; k < 64; k++) hb.stack[k] = h.stack[k];
    for (int k = 0; k < 200; k++) {
        hb.mmI[k] = h.mmI[k];
        hb.mm1n[k] = h.mm1n[k];
        hb.mmM[k] = h.mmM[k];
        ok = ok && fits10(h.mmI[k]) && fits10(h.mm1n[k]);
    }
    for (int k = 0; k < 40; k++) {
        hb.d5[k] = h.d5[k];
        hb.d3[k] = h.d3[k];
    }
    hb.MLbase = M.MLbase;
    hb.MLclosing = M.MLclosing;
    hb.MLintern = M.MLintern;
    hb.TerminalAU = M.TerminalAU;
    hb.bulge1 = M.bulge[1];
    hb.il5_ninio = M.internal_loop[5] + M.ninio;
    for (int k = 0; k < 200; k++) {
        hb.mm23[k] = h.mm23[k];
        hb.mmH[k] = h.mmH[k];
    }
    ok = ok && fits10(M.TerminalAU);
    int ng = 0;
    for (int cls = 0; cls < 3; cls++) {
        hb.cls_begin[cls] = ng;
        for (int u = 0; u <= MAXLOOP; u++) {
            for (int u1 = 0; u1 <= u; u1++) {
                const int k = (u ? h.ncand_upto[u - 1] : 0) + u1;
                const int u2 = u - u1;
                if ((h.cand_code[k] >> 10) != cls) continue;
                const int size = h.cand_size[k] - FB;        // the 10-bit field carries a bias of FB
                ok = ok && size >= -32768 && size < 32768;
                hb.cand[ng++] = (int)(((unsigned)size << 16) | (unsigned)(u1 * WP - u2 + 30));
            }
            hb.ncls_upto[cls][u] = ng - hb.cls_begin[cls];
        }
        hb.ncls_upto[cls][31] = hb.ncls_upto[cls][30];
    }
    hb.cls_begin[3] = ng;
    {
        const int pads[3] = {PADG, PAD1, PADB};
        int o = 0;
        for (int cls = 0; cls < 3; cls++) {
            const int cnt = hb.cls_begin[cls + 1] - hb.cls_begin[cls];
            ok = ok && cnt <= pads[cls];
            for (int k = 0; k < pads[cls]; k++)
                hb.candp[o + k] = k < cnt ? hb.cand[hb.cls_begin[cls] + k] : (int)((unsigned)32767 << 16);   // offset 0, size 32767
            o += pads[cls];
        }
    }
    for (int k = ng; k < NCAND; k++) hb.cand[k] = 0;
    g_mfe4_ok = ok;
    if (!g_dtabB) cudaMalloc(&g_dtabB, sizeof(TabB));
    cudaMemcpy(g_dtabB, &hb, sizeof(TabB), cudaMemcpyHostToDevice);
}

bool mfe4_supports(int n) { return g_mfe4_ok && n >= 2 * BS; }

size_t mfe4_pitch(int n) { return (size_t)((n + BS - 1) / BS) * BS; }

// device bytes one fold of length n needs while it is being computed: three matrices + enforced-pair rows + stack
size_t mfe4_bytes_per_fold(int n) {
    const size_t np = mfe4_pitch(n);
    return 3 * np * np * 4 + 3 * (size_t)(n + 2) * 4 + 3 * (size_t)(n + 8) * 4;
}

void launch_mfe4(const MfeLaunch &L, const MfeTables *d_tab, const int32_t *d_hp_len, void *scratch, size_t scratch_bytes,
                 int32_t *pair32, int n_sm, cudaStream_t stream, int *n_launches) {
    if (L.n_fold <= 0) return;
    const int n = L.W;
    const size_t np = mfe4_pitch(n), mat = np * np * 4, per = mfe4_bytes_per_fold(n);
    int cap = (int)std::min<size_t>((size_t)L.n_fold, scratch_bytes / per);
    if (cap < 1) return;   // the caller sizes the scratch for at least one fold
    static bool cfg = false;
    if (!cfg) {
        cudaFuncSetAttribute(mfe4_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM4_BYTES);
        cudaFuncSetAttribute(mfe4_exterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cfg = true;
    }
    const bool want_pairs = L.pair_tbl || pair32;
    for (int f0 = 0; f0 < L.n_fold; f0 += cap) {
        const int nf = std::min(cap, L.n_fold - f0);
        Mfe4Launch A{};
        A.seqs = L.seqs + (size_t)f0 * n;
        A.hc = L.hc ? L.hc + (size_t)f0 * n : nullptr;
        A.sc = L.sc ? L.sc + (size_t)f0 * (n + 1) : nullptr;
        A.hp_len = d_hp_len;
        A.n_fold = nf;
        A.n = n;
        A.NP = (int)np;
        A.NB = (int)(np / BS);
        A.max_span = L.max_span;
        char *p = static_cast<char *>(scratch);
        A.C = reinterpret_cast<int32_t *>(p);
        p += (size_t)cap * mat;
        A.M = reinterpret_cast<int32_t *>(p);
        p += (size_t)cap * mat;
        A.D = reinterpret_cast<int32_t *>(p);
        p += (size_t)cap * mat;
        int32_t *enf = reinterpret_cast<int32_t *>(p);
        p += (size_t)cap * 3 * (n + 2) * 4;
        A.tb_stack = want_pairs ? reinterpret_cast<int32_t *>(p) : nullptr;
        A.e_out = L.e_out + f0;
        A.pair16 = L.pair_tbl ? L.pair_tbl + (size_t)f0 * n : nullptr;
        A.pair32 = pair32 ? pair32 + (size_t)f0 * n : nullptr;
        if (L.hc && !L.hc_simple) {
            enforced_kernel<<<(nf + 63) / 64, 64, 0, stream>>>(A.hc, nf, n, enf);
            A.enf = enf;
            if (n_launches) (*n_launches)++;
        }
        for (int delta = 0; delta < A.NB; delta++) {
            const long long tasks = (long long)nf * (A.NB - delta);
            const int grid = (int)std::min<long long>(tasks, (long long)n_sm * 4 * 4);
            mfe4_block_kernel<<<grid, NT4, SMEM4_BYTES, stream>>>(A, d_tab, g_dtab4, g_dtabB, delta);
            if (n_launches) (*n_launches)++;
        }
        const size_t smx = (size_t)TAB4_INTS * 4 + 32 + (size_t)(n + 2) * 4;
        mfe4_exterior_kernel<<<std::min(nf, n_sm * 8), NT4B, smx, stream>>>(A, d_tab, g_dtab4);
        if (n_launches) (*n_launches)++;
    }
}

}  // namespace sfb
