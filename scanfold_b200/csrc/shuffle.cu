// Window gathering and background shuffles (Philox4x32-10).
//
// Replaces scramble() (ScanFoldFunctions.py:834-851): mono = random.sample permutation (:800-802),
// di = Altschul-Erikson dinucleotide shuffle (:155-277).  The reference shuffles are unseeded, so the
// requirement is distributional equivalence; parity mode feeds host shuffles instead.  Counters are
// (absolute window index, shuffle index, block), so output is independent of launch geometry / GPU count.
#include "device_common.cuh"

namespace sfb {
namespace {

struct Philox {
    uint32_t c0, c1, c2, c3, k0, k1;
    uint32_t buf[4];
    int have;
    __device__ Philox(unsigned long long seed, uint32_t a, uint32_t b, uint32_t tag)
        : c0(0), c1(a), c2(b), c3(tag), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), have(0) {}
    __device__ void refill() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            uint32_t n0 = hi1 ^ x1 ^ ka, n1 = lo1, n2 = hi0 ^ x3 ^ kb, n3 = lo0;
            x0 = n0; x1 = n1; x2 = n2; x3 = n3;
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        buf[0] = x0; buf[1] = x1; buf[2] = x2; buf[3] = x3;
        have = 4;
        c0++;
    }
    __device__ uint32_t next() {
        if (!have) refill();
        return buf[--have];
    }
    // unbiased integer in [0, n)  (Lemire's multiply-shift with rejection)
    __device__ uint32_t below(uint32_t n) {
        uint64_t m = (uint64_t)next() * n;
        uint32_t l = (uint32_t)m;
        if (l < n) {
            uint32_t t = (0u - n) % n;
            while (l < t) {
                m = (uint64_t)next() * n;
                l = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
};

constexpr int SH_MAXW = MAX_W;

__device__ int window_start(int slot, int first_window, int n_windows, int final_slot, int step, int L, int W) {
    if (final_slot && slot == n_windows - 1) return L - W;  // Q5: final-window set uses seq[L-W:L]
    return (first_window + slot) * step;
}

__global__ void gather_kernel(const uint8_t *seq, int L, int W, int step, int first_window, int n_windows,
                              int final_slot, uint8_t *out) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_windows * W) return;
    int slot = (int)(idx / W), k = (int)(idx % W);
    out[idx] = seq[window_start(slot, first_window, n_windows, final_slot, step, L, W) + k];
}

// Reactivity slice of ScanFold.py:523 handed to a 1-based API (Q7): window position p reads the global
// 1-based position start1 + p; position W falls outside the slice and contributes nothing.
__global__ void slice_sc_kernel(const int32_t *es1, int L, int W, int step, int first_window, int n_windows,
                                int final_slot, int32_t *out) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_windows * (W + 1)) return;
    int slot = (int)(idx / (W + 1)), p = (int)(idx % (W + 1));
    int start0 = window_start(slot, first_window, n_windows, final_slot, step, L, W);
    int v = 0;
    if (p >= 1 && p <= W - 1) {
        int g = start0 + 1 + p;  // global 1-based index
        if (g <= L) v = es1[g];
    }
    out[idx] = v;
}

// one thread per shuffle
__global__ void shuffle_kernel(ShuffleLaunch A) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)A.n_windows * A.r) return;
    const int slot = (int)(idx / A.r), sh = (int)(idx % A.r);
    const int W = A.W;
    const int start = window_start(slot, A.first_window, A.n_windows, A.final_slot, A.step, A.L, W);
    const bool is_final = A.final_slot && slot == A.n_windows - 1;
    // the final-window set gets its own counter so it never repeats a regular window's stream
    const uint32_t wctr = is_final ? 0xFFFFFFFFu : (uint32_t)(A.global_window_base + slot);
    Philox rng(A.seed, wctr, (uint32_t)sh, (uint32_t)A.type + 1u);
    const uint8_t *src = A.seq_codes + start;
    uint8_t *dst = A.out + idx * W;

    if (A.type == SHUFFLE_MONO) {
        // Fisher-Yates: uniform over all W! orderings, like random.sample(frag, len(frag))
        for (int k = 0; k < W; k++) dst[k] = src[k];
        for (int k = W - 1; k > 0; k--) {
            uint32_t j = rng.below((uint32_t)k + 1);
            uint8_t t = dst[k];
            dst[k] = dst[j];
            dst[j] = t;
        }
        return;
    }
    // Altschul-Erikson: random Eulerian walk preserving dinucleotide counts and both end nucleotides
    if (W < 3) {
        for (int k = 0; k < W; k++) dst[k] = src[k];
        return;
    }
    int cnt[5][5];
    int outdeg[5], present[5];
    for (int a = 0; a < 5; a++) {
        outdeg[a] = 0;
        present[a] = 0;
        for (int b = 0; b < 5; b++) cnt[a][b] = 0;
    }
    for (int k = 0; k < W; k++) present[src[k]] = 1;
    for (int k = 0; k + 1 < W; k++) {
        cnt[src[k]][src[k + 1]]++;
        outdeg[src[k]]++;
    }
    const int first = src[0], last = src[W - 1];
    int lastedge[5];
    // choose a last edge per vertex until every vertex reaches `last` through last edges
    for (;;) {
        for (int a = 0; a < 5; a++) {
            lastedge[a] = -1;
            if (!present[a] || a == last) continue;
            uint32_t z = rng.below((uint32_t)outdeg[a]);
            int b = 0;
            uint32_t acc = (uint32_t)cnt[a][0];
            while (z >= acc) acc += (uint32_t)cnt[a][++b];
            lastedge[a] = b;
        }
        bool ok = true;
        for (int a = 0; a < 5 && ok; a++) {
            if (!present[a] || a == last) continue;
            int v = a, hops = 0;
            while (v != last && hops < 6) {
                v = lastedge[v];
                hops++;
            }
            if (v != last) ok = false;
        }
        if (ok) break;
    }
    // successor lists in sequence order, stored contiguously per vertex in dst-sized scratch
    uint8_t lst[SH_MAXW];
    int off[5], fill[5], len[5];
    {
        int o = 0;
        for (int a = 0; a < 5; a++) {
            off[a] = o;
            fill[a] = 0;
            len[a] = outdeg[a];
            o += outdeg[a];
        }
    }
    for (int k = 0; k + 1 < W; k++) {
        int a = src[k];
        lst[off[a] + fill[a]++] = src[k + 1];
    }
    for (int a = 0; a < 5; a++) {
        if (lastedge[a] < 0) continue;
        // remove the first occurrence of the last edge, shuffle the rest, put the last edge at the end
        int n = len[a], pos = 0;
        while (lst[off[a] + pos] != lastedge[a]) pos++;
        for (int k = pos; k + 1 < n; k++) lst[off[a] + k] = lst[off[a] + k + 1];
        n--;
        for (int k = n - 1; k > 0; k--) {
            uint32_t j = rng.below((uint32_t)k + 1);
            uint8_t t = lst[off[a] + k];
            lst[off[a] + k] = lst[off[a] + j];
            lst[off[a] + j] = t;
        }
        lst[off[a] + n] = (uint8_t)lastedge[a];
    }
    if (lastedge[last] < 0 && len[last] > 1) {  // the end vertex has no forced last edge: shuffle all its edges
        int n = len[last];
        for (int k = n - 1; k > 0; k--) {
            uint32_t j = rng.below((uint32_t)k + 1);
            uint8_t t = lst[off[last] + k];
            lst[off[last] + k] = lst[off[last] + j];
            lst[off[last] + j] = t;
        }
    }
    int cur[5] = {0, 0, 0, 0, 0};
    int prev = first;
    dst[0] = (uint8_t)first;
    for (int k = 1; k < W - 1; k++) {
        int ch = lst[off[prev] + cur[prev]++];
        dst[k] = (uint8_t)ch;
        prev = ch;
    }
    dst[W - 1] = (uint8_t)last;
}

}  // namespace

void launch_gather_windows(const uint8_t *seq_codes, int L, int W, int step, int first_window, int n_windows,
                           int final_slot, uint8_t *out, cudaStream_t stream, int *n_launches) {
    long long n = (long long)n_windows * W;
    if (n <= 0) return;
    gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seq_codes, L, W, step, first_window, n_windows,
                                                                  final_slot, out);
    if (n_launches) (*n_launches)++;
}

void launch_slice_hc(const uint8_t *hc, int L, int W, int step, int first_window, int n_windows, int final_slot,
                     uint8_t *out, cudaStream_t stream, int *n_launches) {
    launch_gather_windows(hc, L, W, step, first_window, n_windows, final_slot, out, stream, n_launches);
}

void launch_slice_sc(const int32_t *es1, int L, int W, int step, int first_window, int n_windows, int final_slot,
                     int32_t *out, cudaStream_t stream, int *n_launches) {
    long long n = (long long)n_windows * (W + 1);
    if (n <= 0) return;
    slice_sc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(es1, L, W, step, first_window, n_windows,
                                                                    final_slot, out);
    if (n_launches) (*n_launches)++;
}

void launch_shuffle(const ShuffleLaunch &A, cudaStream_t stream, int *n_launches) {
    long long n = (long long)A.n_windows * A.r;
    if (n <= 0) return;
    shuffle_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(A);
    if (n_launches) (*n_launches)++;
}

}  // namespace sfb
