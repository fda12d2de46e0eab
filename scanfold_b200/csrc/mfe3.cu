// MFE fold kernel, third generation: one CTA per fold, energy only, no constraints, windows up to 320 nt.
//
// Replaces the r background folds per window of energies()/rna_folder (ScanFoldFunctions.py:774-789,805-814)
// -- more than 99 % of all fold arithmetic of a scan.
//
// What changed against mfe2.cu: the <= 496 interior-loop candidates of a cell are no longer walked one by one.
// For generic loops (both sides >= 2 unpaired, not 2x2 / 2x3) of total size u the energy is
//     internal_loop[u] + min(MAX_NINIO, |u1-u2| * ninio) + mismatchI(outer) + mismatchI(inner),
// and all candidates of one u lie on ONE inner diagonal dd = d-2-u at consecutive positions.  With
// G = C + mismatchI(inner) stored per diagonal row, the near-symmetric candidates (|u1-u2| <= 4) are a fixed
// 5-tap (u even) or 4-tap (u odd) weighted stencil along the row, and every other candidate carries the same
// capped asymmetry term, so it only needs a range minimum of G over the row -- including the near ones in that
// range is harmless because the cap is an upper bound of their true term.  Per finished diagonal row the kernel
// stores the two stencils (NE, NO) and a sliding 8-minimum (M8); a cell then needs 1 + (1..4) loads per u
// instead of u-3.  Bulge and 1xn loops keep one load per candidate (rows RB, R1).
//
// Work split: the CTA walks the anti-diagonals two at a time.  Phase C: the pairable cells of both diagonals
// are compacted into lists and every (list chunk, candidate class) pair is an independent work item for one
// warp (classes: generic / table-driven shapes + hairpin + multiloop closing / left bulge + 1xn / right bulge +
// 1xn); each item leaves a partial minimum per cell.  Phase S: per row element the partial minima are combined
// into C and the derived rows G, R1, RB, NE, NO, M8 are written (warp shuffles along the row).  Phase M: the
// multiloop matrix FML of both diagonals, sharing the left operand of the split loop.
// int16 storage is exact as long as no stored energy drops below LOW16; a fold that does is flagged and redone
// by the int32 kernel (mfe.cu) in the same stream, so results never depend on which kernel ran.
#include <cstdlib>
#include <type_traits>

#include "device_common.cuh"

namespace sfb {
namespace {

constexpr int INF16 = 16000;
constexpr int FIN16 = 4000;
constexpr int LOW16 = -12000;
constexpr int R32 = 32;  // ring depth of the rows read by loops of size <= 30 (written after the diagonal pair)
constexpr int R16 = 16;  // ring depth of raw C / pair type / G (read for small loops only)
constexpr int SEG = 25;  // row elements finished per warp pass in phase S (32 lanes - 7 halo lanes of the 8-minimum)

struct alignas(16) Tab3 {
    short stack[64], mmI[200], mm1n[200], mm23[200], mmH[200];
    short mlclose[200];        // mismatchM + TerminalAU + MLintern + MLclosing (closing pair of a multiloop)
    short mlstem[8 * 36];      // [type][5' code][3' code], code 5 = no neighbour: E_MLstem
    short ext[8 * 36];         // E_ExtLoop likewise
    short tAU[8];
    short bulge1;              // bulge[1]
    short il5_ninio;           // internal_loop[5] + ninio   (2x3 loops)
    short MLbase;
    short w1, w2, w3, w4;      // min(MAX_NINIO, k * ninio)
    short pad;
    unsigned char ptype[36];   // pair type of codes a*6+b (code 5 = sentinel)
    unsigned char rtype[8];
};

__constant__ int c3_il[32];     // internal_loop[u]                       (stencil rows already carry the asymmetry)
__constant__ int c3_cap[32];    // internal_loop[u] + MAX_NINIO
__constant__ int c3_size1[32];  // 1xn loops of total size u
__constant__ int c3_sizeB[32];  // bulge[u]
__constant__ int c3_sG6[4];     // generic loops of size 6: (2,4) (3,3) (4,2)
Tab3 *g_dtab3 = nullptr;
bool g_mfe3_ok = false;

template <int P>
struct Smem3 {
    Tab3 tb;
    // ring rows, INF-initialised per fold; ne..rb double as the staging area of C for the exterior loop
    short ne[R32 * P], no[R32 * P], m8[R32 * P], r1[R32 * P], rb[R32 * P];
    short g[R16 * P];
    short rc[R16 * P];
    short dml[4 * P];
    short fml[(P / 2 + 1) * P + 8];
    short part[4 * 2 * P];
    short f5[P + 8];
    unsigned char ctx[R16 * P];
    unsigned char list[2 * P];
    unsigned char sx[P + 8];   // sx[k+1] = code of nucleotide k, sx[0] = sx[W+1] = 5
    int cnt[2];
    int minv[32];
};

__host__ __device__ __forceinline__ int tri4(int d, int W) {  // first cell of diagonal d in the d >= 4 triangle
    return (d - 4) * W - ((d - 1) * d / 2 - 6);
}

template <int A, int B, class F>
__device__ __forceinline__ void sfor(F &&f) {
    if constexpr (A <= B) {
        f(std::integral_constant<int, A>{});
        sfor<A + 1, B>(f);
    }
}

__device__ int hairpin_special3(const MfeTables *T, const Tab3 &tb, const unsigned char *sx, int i, int j, int type) {
    // loops of 3, 4 and 6 nucleotides: tabulated tri- / tetra- / hexaloops; sx is offset by one
    const int u = j - i - 1;
    int e = T->hairpin_len[u];
    auto key = [&](int n) {
        int k = 0, mul = 1;
        for (int t = 0; t < n; t++) {
            k += sx[i + 1 + t] * mul;
            mul *= 5;
        }
        return k;
    };
    if (u == 4) {
        const int k = key(6);
        for (int t = 0; t < T->n_tetra; t++)
            if (T->tetra_key[t] == k) return T->tetra_e[t];
    } else if (u == 6) {
        const int k = key(8);
        for (int t = 0; t < T->n_hexa; t++)
            if (T->hexa_key[t] == k) return T->hexa_e[t];
    } else if (u == 3) {
        const int k = key(5);
        for (int t = 0; t < T->n_tri; t++)
            if (T->tri_key[t] == k) return T->tri_e[t];
        return e + tb.tAU[type];
    }
    return e + tb.mmH[(type * 5 + sx[i + 2]) * 5 + sx[j]];
}

// ---- phase C, class A: generic interior loops of size U (see the header) -------------------------------------
template <int U, int P>
__device__ __forceinline__ void generic_u(const Smem3<P> &sm, int d, int i, int &g0, int &g1) {
    const int dd = d - 2 - U;
    const int r32 = (dd & (R32 - 1)) * P + i, r16 = (dd & (R16 - 1)) * P + i;
    if constexpr (U == 6) {
        g0 = __viaddmin_s32(sm.g[r16 + 3], c3_sG6[0], g0);
        g1 = __viaddmin_s32(sm.g[r16 + 4], c3_sG6[1], g1);
        g0 = __viaddmin_s32(sm.g[r16 + 5], c3_sG6[2], g0);
    } else {
        constexpr int m = U / 2;
        if constexpr (U % 2 == 0)
            g0 = __viaddmin_s32(sm.ne[r32 + 3 + m], c3_il[U], g0);
        else
            g1 = __viaddmin_s32(sm.no[r32 + 3 + m], c3_il[U], g1);
        if constexpr (U == 9 || U == 10) {  // the two capped candidates (u1 = 2 and u2 = 2)
            g0 = __viaddmin_s32(sm.g[r16 + 3], c3_cap[U], g0);
            g1 = __viaddmin_s32(sm.g[r16 + U - 1], c3_cap[U], g1);
        }
        if constexpr (U >= 11) {  // range minimum over u1 = 2 .. U-2, i.e. row positions i+3 .. i+U-1
            int f = sm.m8[r32 + 10];
            if constexpr (U > 11) f = min(f, (int)sm.m8[r32 + U - 1]);
            if constexpr (U > 19) f = min(f, (int)sm.m8[r32 + 18]);
            if constexpr (U > 27) f = min(f, (int)sm.m8[r32 + 26]);
            if constexpr (U % 2 == 0)
                g1 = __viaddmin_s32(f, c3_cap[U], g1);
            else
                g0 = __viaddmin_s32(f, c3_cap[U], g0);
        }
    }
}

// ---- phase C, classes L / R: bulges and 1xn loops with the unpaired stretch on the 3' (L) or 5' (R) side --------
template <int U, int P, bool RIGHT>
__device__ __forceinline__ void side_u(const Smem3<P> &sm, int d, int i, int &aB, int &a1) {
    const int dd = d - 2 - U;
    const int r32 = (dd & (R32 - 1)) * P + i;
    aB = __viaddmin_s32(sm.rb[r32 + (RIGHT ? 1 + U : 1)], c3_sizeB[U], aB);
    if constexpr (U >= 4) a1 = __viaddmin_s32(sm.r1[r32 + (RIGHT ? U : 2)], c3_size1[U], a1);
}

template <int P, int NW, int OCC>
__global__ void __launch_bounds__(NW * 32, OCC)
mfe3_kernel(MfeLaunch L, const MfeTables *__restrict__ T, const Tab3 *__restrict__ gtab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem3<P> &sm = *reinterpret_cast<Smem3<P> *>(smem_raw);
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned full = 0xffffffffu;
    {
        const int *src = reinterpret_cast<const int *>(gtab);
        int *dst = reinterpret_cast<int *>(&sm.tb);
        static_assert(sizeof(Tab3) % 4 == 0, "Tab3 is copied as 32-bit words");
        for (int k = tid; k < (int)(sizeof(Tab3) / 4); k += NT) dst[k] = src[k];
    }
    const Tab3 &tb = sm.tb;
    const unsigned char *sx = sm.sx;
    const int W = L.W, H = W / 2;
    short *gC = reinterpret_cast<short *>(L.gscratch) + (size_t)blockIdx.x * L.gscratch_per_cta;
    auto fidx = [&](int dd, int idx) { return dd <= H ? dd * P + idx : (W - dd) * P + dd + idx; };

    // compacted list of the pairable cells of diagonal d -> sm.list[slot], sm.cnt[slot] (one warp)
    auto build_list = [&](int slot, int d) {
        int nl = 0;
        const int ncells = W - d;
        for (int i0 = 0; i0 < ncells; i0 += 32) {
            const int i = i0 + lane;
            const int t = i < ncells ? tb.ptype[sx[i + 1] * 6 + sx[i + d + 1]] : 0;
            const unsigned m = __ballot_sync(full, t != 0);
            if (t) sm.list[slot * P + nl + __popc(m & ((1u << lane) - 1))] = (unsigned char)i;
            nl += __popc(m);
        }
        if (lane == 0) sm.cnt[slot] = nl;
    };

    for (int fold = blockIdx.x; fold < L.n_fold; fold += gridDim.x) {
        __syncthreads();
        // ---- prologue: sequence with sentinels, INF in every ring row
        for (int k = tid; k < W + 2; k += NT)
            sm.sx[k] = (k == 0 || k == W + 1) ? 5 : L.seqs[(size_t)fold * W + k - 1];
        {
            const int4 inf4 = make_int4(INF16 * 65537, INF16 * 65537, INF16 * 65537, INF16 * 65537);
            int4 *p = reinterpret_cast<int4 *>(sm.ne);
            constexpr int n16 = (int)((5 * R32 * P + 2 * R16 * P + 4 * P) * sizeof(short) / 16);  // ne .. dml
            static_assert(((5 * R32 * P + 2 * R16 * P + 4 * P) * sizeof(short)) % 16 == 0, "ring area is filled as int4");
            for (int k = tid; k < n16; k += NT) p[k] = inf4;
        }
        int minv = 0;
        __syncthreads();
        if (warp < 2) build_list(warp, TURN + 1 + warp);
        __syncthreads();

        for (int d0 = TURN + 1; d0 < W; d0 += 2) {
            const int nd = d0 + 1 < W ? 2 : 1;
            // =================== phase C: partial minima of the pairable cells =======================
            {
                const int n0 = sm.cnt[0], n1 = nd == 2 ? sm.cnt[1] : 0;
                const int nch0 = (n0 + 31) >> 5, nch = nch0 + ((n1 + 31) >> 5);
                for (int it = warp; it < 4 * nch; it += NW) {
                    const int part = it / nch, ch = it - part * nch;
                    const int ds = ch >= nch0 ? 1 : 0, c = ds ? ch - nch0 : ch;
                    const int d = d0 + ds, n = ds ? n1 : n0;
                    const int idx = c * 32 + lane;
                    const bool active = idx < n;
                    const int i = active ? sm.list[ds * P + idx] : 0;
                    const int j = i + d;
                    const int type = tb.ptype[sx[i + 1] * 6 + sx[j + 1]];
                    const int si1 = sx[i + 2], sj1 = sx[j];
                    const int mi = (type * 5 + si1) * 5 + sj1;
                    const int umax = d - 2 - (TURN + 1);   // largest loop size with an inner diagonal > TURN
                    int res;
                    if (part == 0) {
                        int g0 = INF16, g1 = INF16;
                        if (umax >= 6) sfor<6, 13>([&](auto U) { generic_u<decltype(U)::value, P>(sm, d, i, g0, g1); });
                        if (umax >= 14) sfor<14, 21>([&](auto U) { generic_u<decltype(U)::value, P>(sm, d, i, g0, g1); });
                        if (umax >= 22) sfor<22, 30>([&](auto U) { generic_u<decltype(U)::value, P>(sm, d, i, g0, g1); });
                        res = min(g0, g1) + tb.mmI[mi];
                    } else if (part == 1) {
                        // the nine table-driven shapes, the hairpin and the multiloop closing
                        int aT = INF16;
                        auto inner = [&](int u1, int u2, int &cc, int &t2, int &sp1, int &sq1) {
                            const int dd = d - 2 - u1 - u2, p = i + 1 + u1, q = j - 1 - u2;
                            const bool ok = dd > TURN;
                            cc = ok ? sm.rc[(dd & (R16 - 1)) * P + p] : INF16;
                            t2 = ok ? sm.ctx[(dd & (R16 - 1)) * P + p] : 0;
                            sp1 = sx[p];       // S[p-1]
                            sq1 = sx[q + 2];   // S[q+1]
                        };
                        int cc, t2, sp1, sq1;
                        inner(0, 0, cc, t2, sp1, sq1);
                        aT = min(aT, cc + tb.stack[type * 8 + t2]);
                        inner(0, 1, cc, t2, sp1, sq1);
                        aT = min(aT, cc + tb.bulge1 + tb.stack[type * 8 + t2]);
                        inner(1, 0, cc, t2, sp1, sq1);
                        aT = min(aT, cc + tb.bulge1 + tb.stack[type * 8 + t2]);
                        inner(1, 1, cc, t2, sp1, sq1);
                        aT = min(aT, cc + __ldg(&T->int11[type][t2][si1][sj1]));
                        inner(1, 2, cc, t2, sp1, sq1);
                        aT = min(aT, cc + __ldg(&T->int21[type][t2][si1][sq1][sj1]));
                        inner(2, 1, cc, t2, sp1, sq1);
                        aT = min(aT, cc + __ldg(&T->int21[t2][type][sq1][si1][sp1]));
                        inner(2, 2, cc, t2, sp1, sq1);
                        aT = min(aT, cc + __ldg(&T->int22[type][t2][si1][sp1][sq1][sj1]));
                        inner(2, 3, cc, t2, sp1, sq1);
                        aT = min(aT, cc + tb.il5_ninio + tb.mm23[mi] + tb.mm23[(t2 * 5 + sq1) * 5 + sp1]);
                        inner(3, 2, cc, t2, sp1, sq1);
                        aT = min(aT, cc + tb.il5_ninio + tb.mm23[mi] + tb.mm23[(t2 * 5 + sq1) * 5 + sp1]);
                        int eh = __ldg(&T->hairpin_len[d - 1]) + tb.mmH[mi];
                        if (d <= 7 && active) eh = hairpin_special3(T, tb, sx, i, j, type);
                        const int dm = sm.dml[((d - 2) & 3) * P + i + 1];
                        res = min(min(aT, eh), dm + tb.mlclose[(tb.rtype[type] * 5 + sj1) * 5 + si1]);
                    } else if (part == 2) {
                        int aB = INF16, a1 = INF16, bB = INF16, b1 = INF16;
                        if (umax >= 2) sfor<2, 9>([&](auto U) { side_u<decltype(U)::value, P, false>(sm, d, i, (decltype(U)::value & 1) ? bB : aB, (decltype(U)::value & 1) ? b1 : a1); });
                        if (umax >= 10) sfor<10, 19>([&](auto U) { side_u<decltype(U)::value, P, false>(sm, d, i, (decltype(U)::value & 1) ? bB : aB, (decltype(U)::value & 1) ? b1 : a1); });
                        if (umax >= 20) sfor<20, 30>([&](auto U) { side_u<decltype(U)::value, P, false>(sm, d, i, (decltype(U)::value & 1) ? bB : aB, (decltype(U)::value & 1) ? b1 : a1); });
                        res = min(min(a1, b1) + tb.mm1n[mi], min(aB, bB) + tb.tAU[type]);
                    } else {
                        int aB = INF16, a1 = INF16, bB = INF16, b1 = INF16;
                        if (umax >= 2) sfor<2, 9>([&](auto U) { side_u<decltype(U)::value, P, true>(sm, d, i, (decltype(U)::value & 1) ? bB : aB, (decltype(U)::value & 1) ? b1 : a1); });
                        if (umax >= 10) sfor<10, 19>([&](auto U) { side_u<decltype(U)::value, P, true>(sm, d, i, (decltype(U)::value & 1) ? bB : aB, (decltype(U)::value & 1) ? b1 : a1); });
                        if (umax >= 20) sfor<20, 30>([&](auto U) { side_u<decltype(U)::value, P, true>(sm, d, i, (decltype(U)::value & 1) ? bB : aB, (decltype(U)::value & 1) ? b1 : a1); });
                        res = min(min(a1, b1) + tb.mm1n[mi], min(aB, bB) + tb.tAU[type]);
                    }
                    if (active) sm.part[(part * 2 + ds) * P + i] = (short)min(res, INF16);
                }
            }
            __syncthreads();
            // =================== lists of the next diagonal pair (warps 0, 1) ==========================
            if (warp < 2) build_list(warp, d0 + 2 + warp);
            // =================== phase S: C and the derived rows of diagonals d0, d0+1 =================
            {
                const int nseg0 = (W - d0 + SEG - 1) / SEG, nseg = nseg0 + (nd == 2 ? (W - d0 - 1 + SEG - 1) / SEG : 0);
                for (int un = warp; un < nseg; un += NW) {
                    const int ds = un >= nseg0 ? 1 : 0, sg = ds ? un - nseg0 : un;
                    const int d = d0 + ds, ncells = W - d;
                    const int x = sg * SEG - 7 + lane;
                    const bool valid = x >= 0 && x < ncells;
                    const int i = valid ? x : 0, j = i + d;
                    const int t = valid ? tb.ptype[sx[i + 1] * 6 + sx[j + 1]] : 0;
                    int e = INF16;
                    if (t) {
                        const short *pp = sm.part + ds * P + i;
                        e = min(min((int)pp[0], (int)pp[2 * P]), min((int)pp[4 * P], (int)pp[6 * P]));
                        if (e >= FIN16) e = INF16;
                    }
                    int vg = INF16, v1 = INF16, vb = INF16;
                    const int t2 = tb.rtype[t];
                    if (t && i > 0 && j < W - 1 && e < FIN16) {
                        const int m2 = (t2 * 5 + sx[j + 2]) * 5 + sx[i];
                        vg = e + tb.mmI[m2];
                        v1 = e + tb.mm1n[m2];
                        vb = e + tb.tAU[t2];
                    }
                    const int g1 = __shfl_up_sync(full, vg, 1), g2 = __shfl_up_sync(full, vg, 2);
                    const int g3 = __shfl_up_sync(full, vg, 3), g4 = __shfl_up_sync(full, vg, 4);
                    const int ne = min(g2, min(min(g1, g3) + tb.w2, min(vg, g4) + tb.w4));
                    const int no = min(min(g1, g2) + tb.w1, min(vg, g3) + tb.w3);
                    int m8 = min(vg, g1);
                    m8 = min(m8, __shfl_up_sync(full, m8, 2));
                    m8 = min(m8, __shfl_up_sync(full, m8, 4));
                    if (valid && lane >= 7) {
                        minv = min(minv, e);
                        const int o16 = (d & (R16 - 1)) * P + i, o32 = (d & (R32 - 1)) * P + i;
                        sm.rc[o16] = (short)e;
                        sm.ctx[o16] = (unsigned char)t2;
                        sm.g[o16] = (short)vg;
                        sm.r1[o32] = (short)v1;
                        sm.rb[o32] = (short)vb;
                        sm.ne[o32] = (short)ne;
                        sm.no[o32] = (short)no;
                        sm.m8[o32] = (short)m8;
                        gC[tri4(d, W) + i] = (short)(e < FIN16 ? e + tb.ext[t * 36 + sx[i] * 6 + sx[j + 2]] : INF16);
                    }
                }
            }
            // =================== phase M: multiloop matrix of both diagonals ===========================
            {
                const int d = d0, nc0 = W - d, nc1 = nd == 2 ? W - d - 1 : 0;
                const int khi = d - 2 - TURN;   // last split of diagonal d0; diagonal d0+1 has one more
                const int kB = d - 1 - H;       // right operand of diagonal d0 sits in the low half for k >= kB
                for (int i0 = 0; i0 < nc0; i0 += NT) {
                    const int ir = i0 + tid;
                    const int i = min(ir, nc0 - 1);   // the last cell has no neighbour on d0+1: it reads in-bounds garbage there
                    int a0 = 2 * INF16, a1 = 2 * INF16, b0 = 2 * INF16, b1 = 2 * INF16;
                    int k = TURN + 1;
                    {   // A low (+P), B high (+P), B' high (+P)
                        const int kend = min(kB - 1, khi);
                        const short *pa = sm.fml + k * P + i;
                        const short *pb = sm.fml + (W - d + 1 + k) * P + d + i;   // B'(k) = pb[-P + 1]
                        for (; k + 1 <= kend; k += 2, pa += 2 * P, pb += 2 * P) {
                            const int x0 = pa[0], x1 = pa[P];
                            a0 = __viaddmin_s32(x0, pb[0], a0);
                            b0 = __viaddmin_s32(x0, pb[-P + 1], b0);
                            a1 = __viaddmin_s32(x1, pb[P], a1);
                            b1 = __viaddmin_s32(x1, pb[1], b1);
                        }
                        for (; k <= kend; k++, pa += P, pb += P) {
                            const int x0 = pa[0];
                            a0 = __viaddmin_s32(x0, pb[0], a0);
                            b0 = __viaddmin_s32(x0, pb[-P + 1], b0);
                        }
                    }
                    if (k == kB && k <= khi) {   // A low, B low, B' high
                        const int x0 = sm.fml[k * P + i];
                        a0 = __viaddmin_s32(x0, sm.fml[(d - 1 - k) * P + i + k + 1], a0);
                        b0 = __viaddmin_s32(x0, sm.fml[(W - d + k) * P + d + 1 + i], b0);
                        k++;
                    }
                    {   // A low (+P), B low (-P+1), B' low: B'(k) = pb[P]
                        const int kend = min(H, khi);
                        const short *pa = sm.fml + k * P + i;
                        const short *pb = sm.fml + (d - 1 - k) * P + i + k + 1;
                        for (; k + 1 <= kend; k += 2, pa += 2 * P, pb -= 2 * (P - 1)) {
                            const int x0 = pa[0], x1 = pa[P];
                            a0 = __viaddmin_s32(x0, pb[0], a0);
                            b0 = __viaddmin_s32(x0, pb[P], b0);
                            a1 = __viaddmin_s32(x1, pb[-(P - 1)], a1);
                            b1 = __viaddmin_s32(x1, pb[1], b1);
                        }
                        for (; k <= kend; k++, pa += P, pb -= P - 1) {
                            const int x0 = pa[0];
                            a0 = __viaddmin_s32(x0, pb[0], a0);
                            b0 = __viaddmin_s32(x0, pb[P], b0);
                        }
                    }
                    {   // A high (-P+1), B low (-P+1), B' low
                        const short *pa = sm.fml + (W - k) * P + k + i;
                        const short *pb = sm.fml + (d - 1 - k) * P + i + k + 1;
                        for (; k + 1 <= khi; k += 2, pa -= 2 * (P - 1), pb -= 2 * (P - 1)) {
                            const int x0 = pa[0], x1 = pa[-(P - 1)];
                            a0 = __viaddmin_s32(x0, pb[0], a0);
                            b0 = __viaddmin_s32(x0, pb[P], b0);
                            a1 = __viaddmin_s32(x1, pb[-(P - 1)], a1);
                            b1 = __viaddmin_s32(x1, pb[1], b1);
                        }
                        for (; k <= khi; k++, pa -= P - 1, pb -= P - 1) {
                            const int x0 = pa[0];
                            a0 = __viaddmin_s32(x0, pb[0], a0);
                            b0 = __viaddmin_s32(x0, pb[P], b0);
                        }
                    }
                    // the extra split of diagonal d0+1: k = d0 - 4, right operand on diagonal 4
                    if (nd == 2 && d - (TURN + 1) >= TURN + 1) {
                        const int kx = d - (TURN + 1);
                        b1 = __viaddmin_s32(sm.fml[fidx(kx, i)], sm.fml[fidx(TURN + 1, i + kx + 1)], b1);
                    }
                    if (ir < nc0) {
                        const int j = ir + d;
                        int dec = min(a0, a1);
                        if (dec >= FIN16) dec = INF16;
                        int m = dec;
                        if (d - 1 > TURN) {
                            const short *prev = sm.fml + fidx(d - 1, ir);
                            m = min(m, min((int)prev[0], (int)prev[1]) + tb.MLbase);
                        }
                        const int t = tb.ptype[sx[ir + 1] * 6 + sx[j + 1]];
                        if (t) {
                            const short *pp = sm.part + ir;
                            int e = min(min((int)pp[0], (int)pp[2 * P]), min((int)pp[4 * P], (int)pp[6 * P]));
                            if (e < FIN16) m = min(m, e + tb.mlstem[t * 36 + sx[ir] * 6 + sx[j + 2]]);
                        }
                        if (m >= FIN16) m = INF16;
                        minv = min(minv, m);
                        sm.dml[(d & 3) * P + ir] = (short)dec;
                        sm.fml[fidx(d, ir)] = (short)m;
                    }
                    if (ir < nc1) {   // diagonal d0+1: split minimum and stem term now, the neighbour terms after the barrier
                        const int j = ir + d + 1;
                        int dec = min(b0, b1);
                        if (dec >= FIN16) dec = INF16;
                        int m = dec;
                        const int t = tb.ptype[sx[ir + 1] * 6 + sx[j + 1]];
                        if (t) {
                            const short *pp = sm.part + P + ir;
                            int e = min(min((int)pp[0], (int)pp[2 * P]), min((int)pp[4 * P], (int)pp[6 * P]));
                            if (e < FIN16) m = min(m, e + tb.mlstem[t * 36 + sx[ir] * 6 + sx[j + 2]]);
                        }
                        if (m >= FIN16) m = INF16;
                        sm.dml[((d + 1) & 3) * P + ir] = (short)dec;
                        sm.fml[fidx(d + 1, ir)] = (short)m;
                    }
                }
            }
            __syncthreads();
            if (nd == 2) {
                const int d = d0 + 1;
                for (int ir = tid; ir < W - d; ir += NT) {
                    const short *prev = sm.fml + fidx(d - 1, ir);
                    int m = sm.fml[fidx(d, ir)];
                    m = min(m, min((int)prev[0], (int)prev[1]) + tb.MLbase);
                    if (m >= FIN16) m = INF16;
                    minv = min(minv, m);
                    sm.fml[fidx(d, ir)] = (short)m;
                }
            }
            __syncthreads();
        }

        // ---- exterior loop: stage C (+ stem term) back from the scratch row, then F5 sequentially (warp 0)
        {
            short *cx = sm.ne;
            const int ntri = tri4(W, W);
            for (int k = tid; k < ntri; k += NT) cx[k] = __ldcg(gC + k);
            for (int k = tid; k <= min(W, TURN + 1); k += NT) sm.f5[k] = 0;
            minv = __reduce_min_sync(full, minv);
            if (lane == 0) sm.minv[warp] = minv;
            __syncthreads();
            if (warp == 0) {
                for (int len = TURN + 2; len <= W; len++) {
                    const int j = len - 1;
                    int best = INF16;
                    for (int i = lane; i <= j - TURN - 1; i += 32) best = min(best, sm.f5[i] + cx[tri4(j - i, W) + i]);
                    best = __reduce_min_sync(full, best);
                    if (lane == 0) sm.f5[len] = (short)min((int)sm.f5[len - 1], best);
                    __syncwarp();
                }
                if (lane == 0) {
                    int mv = sm.minv[0];
                    for (int q = 1; q < NW; q++) mv = min(mv, sm.minv[q]);
                    L.e_out[fold] = mv < LOW16 ? MFE_REDO : (int)sm.f5[W];
                }
            }
        }
    }
}

template <int P, int NW, int OCC>
void launch_mfe3_t(const MfeLaunch &L, const MfeTables *d_tab, int n_sm, cudaStream_t stream) {
    const size_t smem = sizeof(Smem3<P>);
    static bool cfg = false;
    if (!cfg) {
        cudaFuncSetAttribute(mfe3_kernel<P, NW, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cfg = true;
    }
    const int grid = L.n_fold < n_sm * OCC ? L.n_fold : n_sm * OCC;
    mfe3_kernel<P, NW, OCC><<<grid, NW * 32, smem, stream>>>(L, d_tab, g_dtab3);
}

}  // namespace

bool mfe3_supports(int W) { return g_mfe3_ok && W >= 16 && W <= 120; }

// scratch rows (one per resident CTA, the most any configuration launches)
int mfe3_max_ctas(int n_sm) { return n_sm * 6; }

void mfe3_upload_tables(const MfeTables &M) {
    static Tab3 h;
    auto mm = [](int t, int a, int b) { return (t * 5 + a) * 5 + b; };
    static const int rt[8] = {0, 2, 1, 4, 3, 6, 5, 7};
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++) h.stack[a * 8 + b] = (short)M.stack[a][b];
    for (int t = 0; t < 8; t++) {
        h.tAU[t] = (short)(t > 2 ? M.TerminalAU : 0);
        h.rtype[t] = (unsigned char)rt[t];
        for (int a = 0; a < 5; a++)
            for (int b = 0; b < 5; b++) {
                h.mmI[mm(t, a, b)] = (short)M.mismatchI[t][a][b];
                h.mm1n[mm(t, a, b)] = (short)M.mismatch1nI[t][a][b];
                h.mm23[mm(t, a, b)] = (short)M.mismatch23I[t][a][b];
                h.mmH[mm(t, a, b)] = (short)M.mismatchH[t][a][b];
                h.mlclose[mm(t, a, b)] = (short)(M.mismatchM[t][a][b] + (t > 2 ? M.TerminalAU : 0) + M.MLintern + M.MLclosing);
            }
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) {
                int em = 0, ee = 0;
                if (a < 5 && b < 5) {
                    em = M.mismatchM[t][a][b];
                    ee = M.mismatchExt[t][a][b];
                } else if (a < 5) {
                    em = ee = M.dangle5[t][a];
                } else if (b < 5) {
                    em = ee = M.dangle3[t][b];
                }
                const int au = t > 2 ? M.TerminalAU : 0;
                h.mlstem[t * 36 + a * 6 + b] = (short)(em + au + M.MLintern);
                h.ext[t * 36 + a * 6 + b] = (short)(ee + au);
            }
    }
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) h.ptype[a * 6 + b] = (unsigned char)((a < 5 && b < 5) ? pair_type(a, b) : 0);
    h.bulge1 = (short)M.bulge[1];
    h.il5_ninio = (short)(M.internal_loop[5] + M.ninio);
    h.MLbase = (short)M.MLbase;
    auto wk = [&](int k) { return (short)std::min(M.max_ninio, k * M.ninio); };
    h.w1 = wk(1);
    h.w2 = wk(2);
    h.w3 = wk(3);
    h.w4 = wk(4);
    h.pad = 0;
    // the stencil / range-minimum split needs every asymmetry >= 5 to be capped and sane magnitudes
    g_mfe3_ok = M.ninio >= 0 && 5 * M.ninio >= M.max_ninio && M.max_ninio >= 0 && M.max_ninio < 2000;
    static int sI[32], s1[32], sB[32], sC[32], s6[4];
    for (int u = 0; u < 32; u++) {
        const bool ok = u <= MAXLOOP && M.internal_loop[u < 31 ? u : 30] < INF;
        sI[u] = ok ? M.internal_loop[u] : INF16;
        sC[u] = ok ? M.internal_loop[u] + M.max_ninio : INF16;
        s1[u] = ok ? M.internal_loop[u] + std::min(M.max_ninio, (u - 2) * M.ninio) : INF16;
        sB[u] = (u <= MAXLOOP && M.bulge[u < 31 ? u : 30] < INF) ? M.bulge[u] : INF16;
    }
    for (int k = 0; k < 3; k++) {
        const int u1 = 2 + k, u2 = 6 - u1, diff = u1 > u2 ? u1 - u2 : u2 - u1;
        s6[k] = M.internal_loop[6] < INF ? M.internal_loop[6] + std::min(M.max_ninio, diff * M.ninio) : INF16;
    }
    s6[3] = INF16;
    cudaMemcpyToSymbol(c3_il, sI, sizeof(sI));
    cudaMemcpyToSymbol(c3_cap, sC, sizeof(sC));
    cudaMemcpyToSymbol(c3_size1, s1, sizeof(s1));
    cudaMemcpyToSymbol(c3_sizeB, sB, sizeof(sB));
    cudaMemcpyToSymbol(c3_sG6, s6, sizeof(s6));
    if (!g_dtab3) cudaMalloc(&g_dtab3, sizeof(Tab3));
    cudaMemcpy(g_dtab3, &h, sizeof(Tab3), cudaMemcpyHostToDevice);
}

void launch_mfe3(const MfeLaunch &L, const MfeTables *d_tab, int n_sm, cudaStream_t stream, int *n_launches) {
    if (L.n_fold <= 0) return;
    if (L.W <= 64)
        launch_mfe3_t<64, 4, 6>(L, d_tab, n_sm, stream);
    else
        launch_mfe3_t<120, 4, 3>(L, d_tab, n_sm, stream);
    if (n_launches) (*n_launches)++;
}

}  // namespace sfb
