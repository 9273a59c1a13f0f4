// afsk_tx.cu — transmitter path of libafsk_b200.so (sm_100a).
//
// Replaces Transmitter.__getFrames (afskmodem.py:452-469: __bytesToBits :446-450, ECC.encode
// :166-175, training cycles :457-458, terminator :460-462, one tone per coded bit :463-467,
// 4800 zero frames :468) followed by SoundOutput.__convertFrames (:239-244, out[n] =
// frames[n & ~1], odd trailing frame dropped) for B payloads.  Elementwise and write-bound:
// each thread produces 16-byte vectors (8 samples) with coalesced 128-bit stores; the frame
// value is recomputed from the payload byte (Hamming(7,4) codeword bit -> mark/space square
// wave phase), nothing is staged in HBM.
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "afsk_common.cuh"

namespace {

constexpr int kSynthThreads = 256;
constexpr int kVecPerThread = 32;
constexpr int kChunkVecs = kSynthThreads * kVecPerThread;   // 8192 vectors = 65536 samples per CTA
constexpr int kMaxSymbols = kChunkVecs * 8 / 4 + 8;         // bit symbols one chunk can touch (bf >= 4)

struct __align__(16) TxDesc {
    int64_t pay_off;      // byte offset of the payload
    int64_t pay_len;
    int64_t out_off;      // sample offset of the capture (multiple of 8)
    int64_t out_len;      // frames written by save(): (len(frames) & ~1)
    int64_t total_bits;   // training + terminator + coded bits
    int64_t ts_bits;      // 2 * ts_cycles training bits (1,0,1,0,...)
    int32_t bf;           // = space tone length
    int32_t ml;           // mark tone length: bf, or bf - 2 when bf % 4 == 2 (4800, 960, 1600, 8000, 24000 baud)
    int64_t chunk_first;  // first CTA chunk of this capture
    int64_t nib_off;      // ml != bf: offset of the capture's nibble start table (2 * pay_len + 1 entries)
    int64_t pad;
};

// Hamming(7,4) generator (ECC.__M_GENERATOR :115-123): codeword bit r of nibble v, LSB = c0
__device__ __forceinline__ uint32_t hamming74_encode(uint32_t v)
{
    const uint32_t d0 = (v >> 3) & 1u, d1 = (v >> 2) & 1u, d2 = (v >> 1) & 1u, d3 = v & 1u;
    return (d0 ^ d1 ^ d3) | ((d0 ^ d2 ^ d3) << 1) | (d0 << 2) | ((d1 ^ d2 ^ d3) << 3) | (d1 << 4) | (d2 << 5) |
           (d3 << 6);
}

// bit b of the transmitted sequence (__getFrames :452-469)
__device__ __forceinline__ uint32_t tx_bit(long long b, const TxDesc &d, const uint8_t *__restrict__ pay)
{
    if (b < d.ts_bits) return (uint32_t)(~b & 1);            // training cycle = mark, space  :457-458
    const long long t = b - d.ts_bits;
    if (t < 4) return t == 0 ? 1u : 0u;                      // terminator mark, space x3     :460-462
    const long long j = t - 4;                               // coded bit index
    long long g;                                             // nibble index, high nibble first :446-450
    uint32_t r;
    if (j < 0x7FFFFFFFLL) {                                  // 32-bit divide for all but > 150 MB payloads
        const uint32_t g32 = (uint32_t)j / 7u;
        r = (uint32_t)j - 7u * g32;
        g = g32;
    } else {
        g = j / 7;
        r = (uint32_t)(j - 7 * g);
    }
    const uint32_t byte = pay[d.pay_off + (g >> 1)];
    const uint32_t nib = (g & 1) ? (byte & 15u) : (byte >> 4);
    return (hamming74_encode(nib) >> r) & 1u;
}

// frame value at phase ph of a tone (Waveforms.getSpaceTone/getMarkTone :68-85), as 2 duplicated
// int16 (SoundOutput.__convertFrames :239-244 emits every even frame twice); sym 2 = silence
__device__ __forceinline__ uint32_t tone_pair(uint32_t sym, int ph, int bf)
{
    if (sym == 2u) return 0u;
    const int q = bf >> 2, h = bf >> 1;
    bool hi;
    if (sym) hi = (ph < q) || (ph >= 2 * q && ph < 3 * q);   // mark : q HI, q LO, q HI, q LO  (q = int(bf / 4))
    else hi = ph < h;                                        // space: h HI, h LO
    return hi ? 0x7FFF7FFFu : 0x80008000u;
}

// One CTA writes 8192 consecutive 16-byte vectors of one capture.  Bits are grouped so that a group
// is a whole number of vectors (1 bit when bf % 8 == 0, else 2 bits); the CTA tabulates the vector
// patterns of every symbol combination of a group (3 or 9 combinations: space / mark / silence)
// and the bit symbols its chunk touches in shared memory, then every thread emits
// pattern[combination][vector-in-group] with one 128-bit shared load and one coalesced 128-bit store.
__global__ void __launch_bounds__(kSynthThreads) k_synth(const uint8_t *__restrict__ pay,
                                                         const TxDesc *__restrict__ descs,
                                                         const int32_t *__restrict__ chunk_cap,
                                                         int16_t *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t tx_smem[];
    const int tid = threadIdx.x;
    const long long chunk = blockIdx.x;
    const TxDesc d = descs[chunk_cap[chunk]];
    if (d.ml != d.bf) return;                                // unequal tones: k_synth_var
    const int bf = d.bf;
    const int G = (bf & 7) ? 2 : 1;                          // bits per group
    const int VG = G * bf / 8;                               // vectors per group
    const int ncomb = G == 1 ? 3 : 9;
    uint4 *tab = reinterpret_cast<uint4 *>(tx_smem);
    uint8_t *sym = tx_smem + (size_t)ncomb * VG * 16;

    const long long nvec_cap = (d.out_len + 7) >> 3;
    const long long v0 = (chunk - d.chunk_first) * kChunkVecs;
    const long long vend = min(v0 + (long long)kChunkVecs, nvec_cap);
    const long long g0 = v0 / VG, g1 = (vend - 1) / VG;
    const long long b0 = g0 * G;
    const int nb = (int)(g1 - g0 + 1) * G;
    for (int i = tid; i < nb; i += kSynthThreads) {
        const long long bi = b0 + i;
        sym[i] = bi < d.total_bits ? (uint8_t)tx_bit(bi, d, pay) : (uint8_t)2;    // 4800 zero frames :468
    }
    for (int idx = tid; idx < ncomb * VG; idx += kSynthThreads) {
        const int comb = idx / VG, k = idx - comb * VG;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = 8 * k + 2 * j;                           // even frame inside the group
            uint32_t s = (uint32_t)comb;
            if (G == 2) {
                if (n >= bf) { n -= bf; s = comb % 3; } else s = comb / 3;
            }
            w[j] = tone_pair(s, n, bf);
        }
        tab[idx] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();

    uint4 *dst = reinterpret_cast<uint4 *>(out + d.out_off);
    const int vr0 = (int)(v0 - g0 * VG) + tid;               // vector index relative to group g0
    int gi = vr0 / VG, k = vr0 - gi * VG;
    const int dg = kSynthThreads / VG, dk = kSynthThreads - dg * VG;
#pragma unroll 8
    for (int r = 0; r < kVecPerThread; r++) {
        const long long vi = v0 + r * kSynthThreads + tid;
        if (vi < vend) {
            const int comb = G == 1 ? sym[gi] : sym[2 * gi] * 3 + sym[2 * gi + 1];
            dst[vi] = tab[comb * VG + k];
        }
        gi += dg; k += dk;
        if (k >= VG) { k -= VG; gi++; }
    }
}

// ------------------------------------------------------------------ unequal tone lengths ----
// When bit_frames % 4 == 2 the mark tone (two cycles of int(bf/4) HI + int(bf/4) LO, :81-85) is two
// frames shorter than the space tone, so a bit's first frame depends on how many marks precede it
// (__getFrames :463-467 just appends).  k_tx_nibscan writes the first frame of every Hamming
// codeword (exclusive scan of ones * ml + (7 - ones) * bf over the nibbles), k_synth_var locates
// each output vector's frames by binary search in that table and walks the 7 bits of the codeword.
constexpr int kScanThreads = 256;

__global__ void __launch_bounds__(kScanThreads) k_tx_nibscan(const uint8_t *__restrict__ pay,
                                                            const TxDesc *__restrict__ descs,
                                                            const int32_t *__restrict__ var_caps,
                                                            int64_t *__restrict__ nib_start)
{
    const TxDesc d = descs[var_caps[blockIdx.x]];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ long long warp_tot[kScanThreads / 32];
    int64_t *out = nib_start + d.nib_off;
    const long long G = 2 * d.pay_len;
    long long carry = 0;
    for (long long base = 0; base <= G; base += kScanThreads) {
        const long long g = base + tid;
        long long len = 0;
        if (g < G) {
            const uint32_t byte = pay[d.pay_off + (g >> 1)];
            const int ones = __popc(hamming74_encode((g & 1) ? (byte & 15u) : (byte >> 4)));
            len = (long long)ones * d.ml + (long long)(7 - ones) * d.bf;
        }
        long long inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        long long before = carry, total = 0;
        for (int w = 0; w < kScanThreads / 32; w++) {
            if (w < warp) before += warp_tot[w];
            total += warp_tot[w];
        }
        if (g <= G) out[g] = before + inc - len;             // exclusive: first frame of codeword g
        carry += total;
        __syncthreads();
    }
}

// Position of frame n (even) in the bit sequence of a capture with unequal tones: bit index b (training,
// terminator and coded bits counted together; b >= total_bits: the 4800 zero frames :468) and the frame's
// offset u inside that bit's tone.  One 64-bit division (training) or one binary search over the codeword
// start table (coded bits) — done once per 64 frames by k_synth_var, which then walks bit by bit.
__device__ __forceinline__ void var_locate(long long n, const TxDesc &d, const uint8_t *__restrict__ pay,
                                           const int64_t *__restrict__ ns, long long &b, int &u)
{
    const long long ml = d.ml, sl = d.bf;
    const long long T0 = (d.ts_bits >> 1) * (ml + sl);       // training cycles: mark, space  :457-458
    if (n < T0) {
        const long long cyc = n / (ml + sl);
        const long long r = n - cyc * (ml + sl);
        b = 2 * cyc + (r >= ml ? 1 : 0);
        u = (int)(r >= ml ? r - ml : r);
        return;
    }
    long long t = n - T0;
    if (t < ml) { b = d.ts_bits; u = (int)t; return; }       // terminator: mark, space x3   :460-462
    t -= ml;
    if (t < 3 * sl) { const long long k = t / sl; b = d.ts_bits + 1 + k; u = (int)(t - k * sl); return; }
    t -= 3 * sl;
    const long long G = 2 * d.pay_len;
    if (G == 0 || t >= ns[G]) { b = d.total_bits; u = 0; return; }
    long long lo = 0, hi = G;                                // last codeword starting at or before t
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (ns[mid] <= t) lo = mid; else hi = mid;
    }
    long long rem = t - ns[lo];
    const uint32_t byte = pay[d.pay_off + (lo >> 1)];
    const uint32_t cw = hamming74_encode((lo & 1) ? (byte & 15u) : (byte >> 4));
    int r = 0;
    for (; r < 6; r++) {
        const long long L = ((cw >> r) & 1u) ? ml : sl;
        if (rem < L) break;
        rem -= L;
    }
    b = d.ts_bits + 4 + 7 * lo + r;
    u = (int)rem;
}

// Unequal tones: a lane synthesizes 64 consecutive frames (8 output vectors) into a per-warp staging block in shared
// memory (XOR-swizzled: conflict-free both ways); the warp then writes its 256 vectors as coalesced 128-bit stores.
//
// Tones of at least 8 frames (4800 baud and below): one var_locate, then the lane lists the nine bits its 64 frames can
// touch — symbol and first frame relative to the lane's — in its column of a shared-memory table (inside the coded
// bits the symbols are cut out of one 21-bit stream built from the two payload bytes that hold the three codewords
// involved) and emits its 32 frame pairs WITHOUT A BRANCH: per pair one table load, "has the next bit started" as a
// select, the tone phase as a bit of the tone's mask.  ncu on the walk it replaces (a while loop per pair that
// advanced bit by bit): 14 of 32 lanes active per instruction, 1.04 G warp instructions for 637 M frames, a third of the
// stall samples waiting for instruction fetch — it ran at 0.85 TB/s whatever was done to its searches and its arithmetic.
//
// Shorter tones (8000 baud: 4 / 6 frames; 24000 baud: the mark tone has no frames at all, :81-85) keep that walk: a bit
// list per 64 frames would not be bounded.
constexpr int kVarVecPerLane = 8;
constexpr int kVarBlockVecs = 32 * kVarVecPerLane;           // 256 vectors = 2048 frames per warp step
constexpr int kVarBits = 10;                                 // nine bits per lane step and a sentinel

__global__ void __launch_bounds__(kSynthThreads) k_synth_var(const uint8_t *__restrict__ pay,
                                                             const TxDesc *__restrict__ descs,
                                                             const int32_t *__restrict__ chunk_cap,
                                                             const int64_t *__restrict__ nib_start,
                                                             int16_t *__restrict__ out)
{
    __shared__ uint4 stage[kSynthThreads / 32][kVarBlockVecs];
    __shared__ int32_t bits[kVarBits][kSynthThreads];        // [bit][thread]: (first frame relative to the lane's) * 4 + symbol
    const long long chunk = blockIdx.x;
    const TxDesc d = descs[chunk_cap[chunk]];
    if (d.ml == d.bf) return;                                // equal tones: k_synth
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t *ns = nib_start + d.nib_off;
    const long long nvec_cap = (d.out_len + 7) >> 3;
    const long long v0 = (chunk - d.chunk_first) * kChunkVecs;
    const long long vend = min(v0 + (long long)kChunkVecs, nvec_cap);
    uint4 *dst = reinterpret_cast<uint4 *>(out + d.out_off);
    uint4 *st = stage[warp];
    const bool long_tones = d.ml >= 8;                       // uniform over the CTA
    // tones of up to 64 frames as bit masks over their even phases (bit k: frame 2k is HI; tone_pair's rule, Waveforms :68-85):
    // space = h HI, h LO (h = bf / 2); mark = q HI, q LO, q HI, q LO (q = int(bf / 4))
    const bool use_masks = d.bf <= 64;
    uint32_t mask_mark = 0u, mask_space = 0u;
    if (use_masks) {
        const int q = d.bf >> 2, h = d.bf >> 1;
        auto below = [](int n) -> uint32_t { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); };   // bits k with k < n
        mask_space = below((h + 1) >> 1);                                                // 2k < h
        mask_mark = below((q + 1) >> 1) | (below((3 * q + 1) >> 1) & ~below(q));          // 2k < q  or  2q <= 2k < 3q
    }
    const long long first_coded = d.ts_bits + 4;
    for (long long vb = v0 + (long long)warp * kVarBlockVecs; vb < vend; vb += (long long)(kSynthThreads / 32) * kVarBlockVecs) {
        const long long vl = vb + (long long)lane * kVarVecPerLane;      // this lane's first vector
        if (vl < vend && long_tones) {
            long long b;
            int u;
            var_locate(8 * vl, d, pay, ns, b, u);
            // ---- the nine bits from b: symbols
            uint32_t symbits = 0u;                           // two bits per symbol (2 = the zero frames behind the last bit)
            if (b >= first_coded && b + 8 < d.total_bits) {
                const long long j = b - first_coded;         // coded bit index: codeword g0, place r0
                const long long g0 = j < 0x7FFFFFFFLL ? (long long)((uint32_t)j / 7u) : j / 7;   // 32-bit divide for all but > 150 MB payloads
                const int r0 = (int)(j - 7 * g0);
                const long long i0 = g0 >> 1;
                const uint32_t B0 = pay[d.pay_off + i0];
                const uint32_t B1 = (i0 + 1 < d.pay_len) ? pay[d.pay_off + i0 + 1] : 0u;
                const uint32_t B01 = (B0 << 8) | B1;
                uint32_t stream = 0u;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int idx = (int)(g0 & 1) + k;       // nibble 0..3 of the two bytes, high nibble first :446-450
                    stream |= hamming74_encode((B01 >> (12 - 4 * idx)) & 15u) << (7 * k);
                }
                stream >>= r0;
#pragma unroll
                for (int i = 0; i < 9; i++) symbits |= ((stream >> i) & 1u) << (2 * i);
            } else {
#pragma unroll 1
                for (int i = 0; i < 9; i++) symbits |= (b + i < d.total_bits ? tx_bit(b + i, d, pay) : 2u) << (2 * i);
            }
            // ---- first frames relative to the lane's first frame; bit 0 started u frames before it
            int rel = -u;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const uint32_t sy = (symbits >> (2 * i)) & 3u;
                bits[i][tid] = rel * 4 + (int)sy;
                rel += sy == 2u ? (1 << 20) : (sy ? d.ml : d.bf);
            }
            bits[9][tid] = (1 << 23) + 2;                    // never starts
            // ---- 32 frame pairs: at most one bit starts between two of them (tones of 8 frames and more)
            int cur = 0;
            int curk = bits[0][tid];
#pragma unroll
            for (int k = 0; k < kVarVecPerLane; k++) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int o = 8 * k + 2 * j;
                    const int nextk = bits[cur + 1][tid];
                    const bool adv = o >= (nextk >> 2);
                    cur += adv ? 1 : 0;
                    curk = adv ? nextk : curk;
                    const uint32_t sy = (uint32_t)curk & 3u;
                    const int ph = o - (curk >> 2);
                    if (use_masks)
                        w[j] = sy == 2u ? 0u : ((((sy ? mask_mark : mask_space) >> ((ph >> 1) & 31)) & 1u) ? 0x7FFF7FFFu : 0x80008000u);
                    else
                        w[j] = tone_pair(sy, ph, d.bf);
                }
                const int v = lane * kVarVecPerLane + k;
                st[v ^ ((v >> 3) & 7)] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else if (vl < vend) {
            long long b;
            int u;
            var_locate(8 * vl, d, pay, ns, b, u);
            uint32_t sym = b < d.total_bits ? tx_bit(b, d, pay) : 2u;
            int L = sym == 2u ? 0x7FFFFFFF : (sym ? d.ml : d.bf);
#pragma unroll
            for (int k = 0; k < kVarVecPerLane; k++) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    w[j] = tone_pair(sym, u, d.bf);
                    u += 2;
                    while (u >= L) {                         // next bit of the sequence (__getFrames :463-467 appends tone after tone;
                        u -= L;                              //  at 24000 baud the mark tone has no frames at all, :81-85)
                        b++;
                        sym = b < d.total_bits ? tx_bit(b, d, pay) : 2u;
                        L = sym == 2u ? 0x7FFFFFFF : (sym ? d.ml : d.bf);
                    }
                }
                const int v = lane * kVarVecPerLane + k;
                st[v ^ ((v >> 3) & 7)] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < kVarVecPerLane; r++) {
            const int v = 32 * r + lane;
            if (vb + v < vend) dst[vb + v] = st[v ^ ((v >> 3) & 7)];
        }
        __syncwarp();
    }
}

static size_t synth_smem_bytes(int bf)
{
    const int G = (bf & 7) ? 2 : 1, VG = G * bf / 8;
    return (size_t)(G == 1 ? 3 : 9) * VG * 16 + kMaxSymbols;
}

}  // namespace

struct AfskTxPlan {
    int device = 0;
    int B = 0;
    std::vector<TxDesc> descs;
    std::vector<int64_t> out_off, out_len;
    TxDesc *d_descs = nullptr;
    int32_t *d_chunk_cap = nullptr;
    std::vector<int32_t> chunk_cap;
    int64_t total_chunks = 0;
    size_t smem = 0;
    std::vector<int32_t> var_caps;       // captures with unequal mark/space tone lengths
    int32_t *d_var_caps = nullptr;
    int64_t *d_nib_start = nullptr;
    int64_t nib_entries = 0;
    bool has_equal = false;
};

extern "C" {

int64_t afsk_tx_num_samples(int baud, int64_t ts_cycles, int64_t payload_bytes, const uint8_t *payload)
{
    int bf, ml, sl;
    if (!afsk_tone_geometry(baud, &bf, &ml, &sl)) { afsk_set_error("Invalid baud rate."); return AFSK_E_BAUD; }
    if (ts_cycles < 0) ts_cycles = 0;
    int64_t frames = ts_cycles * (int64_t)(ml + sl) + ml + 3 * (int64_t)sl + AFSK_TAIL_FRAMES;
    if (ml == sl) {
        frames += payload_bytes * 14 * (int64_t)sl;
    } else {
        if (payload_bytes > 0 && !payload) { afsk_set_error("payload needed for unequal tone lengths"); return AFSK_E_ARG; }
        for (int64_t i = 0; i < payload_bytes; i++) {
            for (int half = 0; half < 2; half++) {
                uint32_t v = half == 0 ? payload[i] >> 4 : payload[i] & 15u;
                uint32_t d0 = (v >> 3) & 1u, d1 = (v >> 2) & 1u, d2 = (v >> 1) & 1u, d3 = v & 1u;
                int ones = (int)((d0 ^ d1 ^ d3) + (d0 ^ d2 ^ d3) + d0 + (d1 ^ d2 ^ d3) + d1 + d2 + d3);
                frames += (int64_t)ones * ml + (int64_t)(7 - ones) * sl;
            }
        }
    }
    return frames & ~(int64_t)1;      // SoundOutput.__convertFrames :241 drops an odd trailing frame
}

int afsk_tx_plan_create(int device, int B, const int64_t *h_pay_off, const int32_t *h_baud, const int64_t *h_ts_cycles,
                        const uint8_t *h_payload, AfskTxPlan **plan_out)
{
    if (!plan_out || B < 0 || (B > 0 && (!h_pay_off || !h_baud || !h_ts_cycles))) return AFSK_E_ARG;
    AfskDeviceGuard guard(device);
    if (!guard.ok) { afsk_set_error("cannot select device %d", device); return AFSK_E_CUDA; }
    AfskTxPlan *P = new (std::nothrow) AfskTxPlan();
    if (!P) return AFSK_E_ARG;
    P->device = device;
    P->B = B;
    P->descs.resize(B);
    P->out_off.assign(B + 1, 0);
    P->out_len.assign(B, 0);
    int64_t chunks = 0;
    for (int c = 0; c < B; c++) {
        int bf, ml, sl;
        if (!afsk_tone_geometry(h_baud[c], &bf, &ml, &sl)) {
            delete P; afsk_set_error("Invalid baud rate."); return AFSK_E_BAUD;
        }
        TxDesc &d = P->descs[c];
        const int64_t ts = h_ts_cycles[c] < 0 ? 0 : h_ts_cycles[c];
        d.pay_off = h_pay_off[c];
        d.pay_len = h_pay_off[c + 1] - h_pay_off[c];
        if (d.pay_len < 0) { delete P; afsk_set_error("payload offsets must be non-decreasing"); return AFSK_E_ARG; }
        d.bf = bf; d.ml = ml; d.pad = 0; d.nib_off = 0;
        d.ts_bits = 2 * ts;
        d.total_bits = d.ts_bits + 4 + 14 * d.pay_len;
        if (ml == sl) {
            d.out_len = (d.total_bits * bf + AFSK_TAIL_FRAMES) & ~(int64_t)1;
            P->smem = std::max(P->smem, synth_smem_bytes(bf));
            P->has_equal = true;
        } else {
            // the length depends on how many mark bits the codewords hold: the plan is made for THIS payload
            if (d.pay_len > 0 && !h_payload) {
                delete P; afsk_set_error("afsk_tx_plan_create: the payload is needed for baud %d (unequal tone lengths)", h_baud[c]);
                return AFSK_E_ARG;
            }
            const int64_t nf = afsk_tx_num_samples(h_baud[c], ts, d.pay_len, h_payload ? h_payload + d.pay_off : nullptr);
            if (nf < 0) { delete P; return (int)nf; }
            d.out_len = nf;
            d.nib_off = P->nib_entries;
            P->nib_entries += 2 * d.pay_len + 1;
            P->var_caps.push_back(c);
        }
        d.out_off = P->out_off[c];
        d.chunk_first = chunks;
        const int64_t nvec = (d.out_len + 7) >> 3;
        const int64_t nch = (nvec + kChunkVecs - 1) / kChunkVecs;
        if (chunks + nch > 0x7FFFFFF0LL) { delete P; afsk_set_error("batch too large"); return AFSK_E_ARG; }
        P->chunk_cap.insert(P->chunk_cap.end(), (size_t)nch, (int32_t)c);
        chunks += nch;
        P->out_len[c] = d.out_len;
        P->out_off[c + 1] = P->out_off[c] + nvec * 8;
    }
    P->total_chunks = chunks;
    if (chunks > 0x7FFFFFFFLL) { delete P; afsk_set_error("batch too large"); return AFSK_E_ARG; }
    if (P->smem > 200 * 1024) { delete P; afsk_set_error("bit_frames too large for the synthesizer"); return AFSK_E_UNSUPPORTED; }
    cudaError_t e = cudaMalloc((void **)&P->d_descs, sizeof(TxDesc) * (B ? B : 1));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && B) e = cudaMemcpy(P->d_descs, P->descs.data(), sizeof(TxDesc) * B, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->d_chunk_cap, sizeof(int32_t) * (chunks ? chunks : 1));
    if (e == cudaSuccess && chunks)
        e = cudaMemcpy(P->d_chunk_cap, P->chunk_cap.data(), sizeof(int32_t) * chunks, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !P->var_caps.empty()) {
        e = cudaMalloc((void **)&P->d_var_caps, sizeof(int32_t) * P->var_caps.size());
        if (e == cudaSuccess) e = cudaMemcpy(P->d_var_caps, P->var_caps.data(), sizeof(int32_t) * P->var_caps.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc((void **)&P->d_nib_start, sizeof(int64_t) * P->nib_entries);
    }
    if (e != cudaSuccess) {
        afsk_set_error("afsk_tx_plan_create: %s", cudaGetErrorString(e));
        afsk_tx_plan_destroy(P);
        return AFSK_E_CUDA;
    }
    *plan_out = P;
    return AFSK_OK;
}

int afsk_tx_plan_destroy(AfskTxPlan *P)
{
    if (!P) return AFSK_OK;
    AfskDeviceGuard guard(P->device);
    cudaFree(P->d_descs);
    cudaFree(P->d_chunk_cap);
    cudaFree(P->d_var_caps);
    cudaFree(P->d_nib_start);
    delete P;
    return AFSK_OK;
}

int afsk_tx_plan_out_offsets(const AfskTxPlan *P, const int64_t **h_out_off, const int64_t **h_out_len)
{
    if (!P) return AFSK_E_ARG;
    if (h_out_off) *h_out_off = P->out_off.data();
    if (h_out_len) *h_out_len = P->out_len.data();
    return AFSK_OK;
}

int afsk_tx_synth(AfskTxPlan *P, const uint8_t *d_payload, int16_t *d_out, void *stream)
{
    if (!P || (P->B > 0 && !d_out)) return AFSK_E_ARG;
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) != 0) { afsk_set_error("afsk_tx_synth: d_out must be 16-byte aligned"); return AFSK_E_ARG; }
    if (P->B == 0 || P->total_chunks == 0) return AFSK_OK;
    AfskDeviceGuard guard(P->device);
    if (!guard.ok) return AFSK_E_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    if (P->has_equal)
        k_synth<<<(unsigned)P->total_chunks, kSynthThreads, P->smem, st>>>(d_payload, P->d_descs, P->d_chunk_cap, d_out);
    if (!P->var_caps.empty()) {
        k_tx_nibscan<<<(unsigned)P->var_caps.size(), kScanThreads, 0, st>>>(d_payload, P->d_descs, P->d_var_caps, P->d_nib_start);
        k_synth_var<<<(unsigned)P->total_chunks, kSynthThreads, 0, st>>>(d_payload, P->d_descs, P->d_chunk_cap, P->d_nib_start, d_out);
    }
    AFSK_CUDA(cudaGetLastError());
    return AFSK_OK;
}

int afsk_tx_synth_host(int device, const uint8_t *h_payload, const int64_t *h_pay_off, int B, const int32_t *h_baud,
                       const int64_t *h_ts_cycles, int16_t *h_out, const int64_t *h_out_off)
{
    if (B < 0 || (B > 0 && (!h_pay_off || !h_out || !h_out_off))) return AFSK_E_ARG;
    if (B == 0) return AFSK_OK;
    AfskTxPlan *P = nullptr;
    int rc = afsk_tx_plan_create(device, B, h_pay_off, h_baud, h_ts_cycles, h_payload, &P);
    if (rc) return rc;
    AfskDeviceGuard guard(device);
    const int64_t pay_bytes = h_pay_off[B];
    uint8_t *d_pay = nullptr; int16_t *d_out = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_pay, (size_t)(pay_bytes ? pay_bytes : 16));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, (size_t)(P->out_off[B] ? P->out_off[B] : 8) * 2);
    if (e == cudaSuccess && pay_bytes) e = cudaMemcpy(d_pay, h_payload, (size_t)pay_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = afsk_tx_synth(P, d_pay, d_out, nullptr);
        for (int c = 0; c < B && rc == AFSK_OK && e == cudaSuccess; c++) {
            if (h_out_off[c + 1] - h_out_off[c] < P->out_len[c]) { rc = AFSK_E_ARG; afsk_set_error("output capacity too small for capture %d", c); break; }
            if (P->out_len[c])
                e = cudaMemcpy(h_out + h_out_off[c], d_out + P->out_off[c], (size_t)P->out_len[c] * 2, cudaMemcpyDeviceToHost);
        }
    }
    cudaFree(d_pay); cudaFree(d_out);
    afsk_tx_plan_destroy(P);
    if (e != cudaSuccess) { afsk_set_error("afsk_tx_synth_host: %s", cudaGetErrorString(e)); return AFSK_E_CUDA; }
    return rc;
}

}  // extern "C"
