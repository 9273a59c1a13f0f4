/*
 * afsk_oracle.c — CPU restatement of the reference's RX/TX hot path, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product package (afskmodem_b200) never links, imports or calls it.
 *
 * Parity status: PINNED.  The reference (/root/reference/afskmodem.py) ships no tests or
 * golden vectors of its own, so the pin is the reference itself, executed
 * unmodified in the build container (oracle/ref_harness.py) — tests/golden/*.npz hold its
 * outputs (final bytes / exception, clock index, training-end frame, bit count, byte count,
 * synthesized frames) and tests/test_oracle_golden.py checks every function below against
 * them.  tests/test_oracle_vs_reference.py re-runs the live comparison when the reference is
 * mounted.
 *
 * Every function cites the reference lines it restates.  Loops are kept in the reference's
 * own order (no closed forms) so that this file is an honest "port" CPU baseline as well.
 * All arithmetic is integer; Python's int(x / n) on non-negative x < 2^53 equals x / n in C.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AFSK_RATE 48000
#define HI 32767
#define LO (-32768)

/* status codes shared with tests (mirrors include/afsk_b200.h AFSK_ST_*) */
enum {
    ST_OK = 0,          /* >= 1 coded bit decoded                                  */
    ST_NO_CLOCK = 1,    /* len(frames) < 4096  (afskmodem.py:323-325)               */
    ST_NO_DATA = 2,     /* clock found, 0 bits (no terminator, or quiet at once)    */
    ST_EXC_WAVELEN = -1,/* Exception("Comparing two waveforms of different lengths.") */
    ST_EXC_INDEX = -2,  /* IndexError: scan_diffs[0] with an empty scan (2*bf >= 4096) */
    ST_EXC_BAUD = -3    /* Exception("Invalid baud rate.") from the constructor     */
};

/* ---- Waveforms.getSpaceTone / getMarkTone / getTrainingCycle : afskmodem.py:68-91 ---- */

/* returns length, or -1 for Exception("Invalid baud rate.") */
int afsk_oracle_space_tone(int baud, int16_t *out)
{
    if (baud <= 0 || AFSK_RATE % baud != 0) return -1;           /* :69-70 */
    double bit_frames = (double)AFSK_RATE / (double)baud;        /* :71   */
    int h = (int)(bit_frames / 2.0);                             /* :73   */
    int n = 0;
    for (int i = 0; i < h; i++) out[n++] = HI;                   /* :73-74 */
    for (int i = 0; i < h; i++) out[n++] = LO;                   /* :75-76 */
    return n;
}

int afsk_oracle_mark_tone(int baud, int16_t *out)
{
    if (baud <= 0 || AFSK_RATE % baud != 0) return -1;           /* :81-82 */
    int n = afsk_oracle_space_tone(baud * 2, out);               /* :83   */
    if (n < 0) return -1;
    int n2 = afsk_oracle_space_tone(baud * 2, out + n);          /* :84   */
    return n + n2;
}

int afsk_oracle_training_cycle(int baud, int16_t *out)
{
    int n = afsk_oracle_mark_tone(baud, out);                    /* :89 */
    if (n < 0) return -1;
    int n2 = afsk_oracle_space_tone(baud, out + n);              /* :90 */
    if (n2 < 0) return -1;
    return n + n2;
}

/* ---- Waveforms.getAmplitude : afskmodem.py:94-98 ---- */
int afsk_oracle_amplitude(const int16_t *x, int n)
{
    int64_t sum = 0;
    for (int i = 0; i < n; i++) sum += x[i] < 0 ? -(int64_t)x[i] : (int64_t)x[i];
    return (int)(sum / n);
}

/* ---- Waveforms.getDiff : afskmodem.py:101-107 (lengths are checked by the callers) ---- */
static int get_diff_raw(const int32_t *a, const int16_t *b, int n)     /* b = raw frames */
{
    int64_t total = 0;
    for (int i = 0; i < n; i++) {
        int64_t d = (int64_t)a[i] - (int64_t)b[i];
        total += d < 0 ? -d : d;
    }
    return (int)(total / n);
}

static int get_diff_amp(const int32_t *a, const int32_t *b, int n)     /* b = amplified frames */
{
    int64_t total = 0;
    for (int i = 0; i < n; i++) {
        int64_t d = (int64_t)a[i] - (int64_t)b[i];
        total += d < 0 ? -d : d;
    }
    return (int)(total / n);
}

/* ---- ECC : afskmodem.py:115-175 ---- */
static const uint8_t M_GEN[7][4] = {   /* :115-123 */
    {1, 1, 0, 1}, {1, 0, 1, 1}, {1, 0, 0, 0}, {0, 1, 1, 1}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
static const uint8_t M_PAR[3][7] = {   /* :125-129 */
    {1, 0, 1, 0, 1, 0, 1}, {0, 1, 1, 0, 0, 1, 1}, {0, 0, 0, 1, 1, 1, 1}};

/* ECC.__encodeNibble :141-142 (via __multiply :132-138) */
void afsk_oracle_ecc_encode_nibble(const uint8_t d[4], uint8_t out[7])
{
    for (int i = 0; i < 7; i++) {
        int r = 0;
        for (int j = 0; j < 4; j++) r += M_GEN[i][j] * d[j];
        out[i] = (uint8_t)(r % 2);
    }
}

/* ECC.__decodeNibble :145-151 */
void afsk_oracle_ecc_decode_nibble(const uint8_t c_in[7], uint8_t out[4])
{
    uint8_t c[7], syn[3];
    memcpy(c, c_in, 7);
    for (int i = 0; i < 3; i++) {
        int r = 0;
        for (int j = 0; j < 7; j++) r += M_PAR[i][j] * c[j];
        syn[i] = (uint8_t)(r % 2);
    }
    int error_pos = syn[2] * 4 + syn[1] * 2 + syn[0];            /* :147 */
    if (error_pos != 0) c[error_pos - 1] ^= 1;                   /* :149-150 */
    out[0] = c[2]; out[1] = c[4]; out[2] = c[5]; out[3] = c[6];  /* :151 */
}

/* ECC.decode :154-163 — bits are 0/1 bytes; returns number of decoded bits (4 * (n/7)) */
int64_t afsk_oracle_ecc_decode(const uint8_t *bits, int64_t n, uint8_t *out)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n - 6; i += 7) {                     /* :156 */
        afsk_oracle_ecc_decode_nibble(bits + i, out + m);
        m += 4;
    }
    return m;
}

/* ECC.encode :166-175 — returns number of coded bits (7 * (n/4)) */
int64_t afsk_oracle_ecc_encode(const uint8_t *bits, int64_t n, uint8_t *out)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n - 3; i += 4) {                     /* :168 */
        afsk_oracle_ecc_encode_nibble(bits + i, out + m);
        m += 7;
    }
    return m;
}

/* ---- Receiver ---- */

/* Receiver.__amplify :287-296 */
static void amplify(const int16_t *x, int n, int32_t *out)
{
    for (int i = 0; i < n; i++) out[i] = x[i] > 512 ? HI : (x[i] < -512 ? LO : 0);
}

typedef struct {
    int bf;                 /* Receiver.__bit_frames :277 */
    int mark_len, space_len, train_len;
    int32_t *mark, *space, *train;   /* widened tone tables */
    int32_t *amp;                    /* scratch for amplify */
} rx_ctx;

static int rx_ctx_init(rx_ctx *c, int baud)
{
    memset(c, 0, sizeof(*c));
    if (baud <= 0) return ST_EXC_BAUD;
    c->bf = (int)((double)AFSK_RATE / (double)baud);             /* :277 */
    int cap = 2 * AFSK_RATE + 16;
    int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * cap);
    int ns = afsk_oracle_space_tone(baud, tmp);                  /* :280 */
    if (ns < 0) { free(tmp); return ST_EXC_BAUD; }
    c->space_len = ns;
    c->space = (int32_t *)malloc(sizeof(int32_t) * (ns + 1));
    for (int i = 0; i < ns; i++) c->space[i] = tmp[i];
    int nm = afsk_oracle_mark_tone(baud, tmp);                   /* :281 */
    if (nm < 0) { free(tmp); free(c->space); return ST_EXC_BAUD; }
    c->mark_len = nm;
    c->mark = (int32_t *)malloc(sizeof(int32_t) * (nm + 1));
    for (int i = 0; i < nm; i++) c->mark[i] = tmp[i];
    int nt = afsk_oracle_training_cycle(baud, tmp);              /* :282 */
    c->train_len = nt;
    c->train = (int32_t *)malloc(sizeof(int32_t) * (nt + 1));
    for (int i = 0; i < nt; i++) c->train[i] = tmp[i];
    c->amp = (int32_t *)malloc(sizeof(int32_t) * (c->bf + 1));
    free(tmp);
    return 0;
}

static void rx_ctx_free(rx_ctx *c)
{
    free(c->mark); free(c->space); free(c->train); free(c->amp);
}

/* Receiver.__recoverClockIndex :322-339.  Returns index, -1 (too short), or ST_EXC_* - 100. */
static int64_t recover_clock(const rx_ctx *c, const int16_t *x, int64_t n, int *exc)
{
    *exc = 0;
    if (n < 4096) return -1;                                     /* :323-325 */
    int span = 4096 - c->bf * 2;                                 /* :327 */
    if (span <= 0) { *exc = ST_EXC_INDEX; return -1; }           /* :332 scan_diffs[0] on [] */
    if (c->train_len != c->bf * 2) { *exc = ST_EXC_WAVELEN; return -1; }  /* :102-103 via :329 */
    int min_diff = 0, min_index = 0;
    for (int i = 0; i < span; i++) {                             /* :327-331 */
        int d = get_diff_raw(c->train, x + i, c->bf * 2);
        if (i == 0 || d < min_diff) { min_diff = d; min_index = i; }  /* :332-337 first strict min */
    }
    return min_index;
}

/* Receiver.__decodeBit :342-351 — returns 0/1, or <0 for the getDiff length exception */
static int decode_bit(const rx_ctx *c, const int16_t *chunk)
{
    amplify(chunk, c->bf, c->amp);                               /* :344 */
    if (c->mark_len != c->bf) return ST_EXC_WAVELEN;             /* :346 -> :102-103 */
    int mark_diff = get_diff_amp(c->mark, c->amp, c->bf);        /* :346 */
    if (c->space_len != c->bf) return ST_EXC_WAVELEN;            /* :347 -> :102-103 */
    int space_diff = get_diff_amp(c->space, c->amp, c->bf);      /* :347 */
    return mark_diff < space_diff ? 1 : 0;                       /* :348-351 */
}

typedef struct {
    int32_t status;      /* ST_* */
    int32_t clock;       /* "Recovered clock. (frame N)" :338, -1 if none      */
    int64_t train_end;   /* "Training sequence terminated on frame N" :368     */
    int64_t nbits;       /* "Decoded N bits. (including ECC)" :380             */
    int64_t nbytes;      /* "Decoded N bytes." :427                            */
} afsk_oracle_result;

/*
 * Receiver.__decodeBits :354-381 + load :420-430 (minus the utf-8 step, which stays in Python).
 * coded_bits (optional, capacity >= n/bf + 1) receives the raw 0/1 coded bits;
 * out (capacity >= n/bf/14 + 1) receives the decoded bytes.
 */
int afsk_oracle_rx_decode(const int16_t *x, int64_t n, int baud, int amp_end_threshold,
                          uint8_t *out, uint8_t *coded_bits, afsk_oracle_result *res)
{
    rx_ctx c;
    res->status = ST_NO_CLOCK; res->clock = -1; res->train_end = -1; res->nbits = 0; res->nbytes = 0;
    int rc = rx_ctx_init(&c, baud);
    if (rc) { res->status = rc; return rc; }
    int exc = 0;
    int64_t i = recover_clock(&c, x, n, &exc);                   /* :356 */
    if (exc) { res->status = exc; rx_ctx_free(&c); return exc; }
    if (i == -1) { rx_ctx_free(&c); return ST_NO_CLOCK; }        /* :357-358 */
    res->clock = (int32_t)i;

    int seq[4] = {0, 0, 0, 0};                                   /* :361 */
    while (i < n - c.bf) {                                       /* :362 (strict) */
        const int16_t *chunk = x + i;                            /* :363 */
        i += c.bf;                                               /* :364 */
        int b = decode_bit(&c, chunk);                           /* :365 */
        if (b < 0) { res->status = b; rx_ctx_free(&c); return b; }
        seq[0] = seq[1]; seq[1] = seq[2]; seq[2] = seq[3]; seq[3] = b;   /* :387-389 */
        if (seq[0] == 1 && seq[1] == 0 && seq[2] == 0 && seq[3] == 0) break;  /* :390, :366 */
    }
    res->train_end = i;                                          /* :368 */

    int64_t cap = n / c.bf + 2;
    uint8_t *bits = coded_bits ? coded_bits : (uint8_t *)malloc((size_t)cap);
    int64_t nb = 0;
    while (i < n - c.bf) {                                       /* :372 */
        const int16_t *chunk = x + i;                            /* :373 */
        if (afsk_oracle_amplitude(chunk, c.bf) < amp_end_threshold) break;   /* :375-376 */
        int b = decode_bit(&c, chunk);                           /* :377 */
        if (b < 0) { res->status = b; if (!coded_bits) free(bits); rx_ctx_free(&c); return b; }
        bits[nb++] = (uint8_t)b;
        i += c.bf;                                               /* :378 */
    }
    res->nbits = nb;                                             /* :380 */
    if (nb == 0) {                                               /* load :422-424 "No data." */
        res->status = ST_NO_DATA;
        if (!coded_bits) free(bits);
        rx_ctx_free(&c);
        return ST_NO_DATA;
    }
    uint8_t *dec = (uint8_t *)malloc((size_t)(nb / 7 * 4 + 8));
    int64_t nd = afsk_oracle_ecc_decode(bits, nb, dec);          /* :425 */
    int64_t nbytes = 0;
    for (int64_t k = 0; k <= nd - 8; k += 8) {                   /* __bitsToBytes :393-399 */
        int v = 0;
        for (int j = 0; j < 8; j++) v = (v << 1) | dec[k + j];   /* int(bits, 2): MSB first */
        out[nbytes++] = (uint8_t)v;
    }
    res->nbytes = nbytes;                                        /* :427 */
    res->status = ST_OK;
    free(dec);
    if (!coded_bits) free(bits);
    rx_ctx_free(&c);
    return ST_OK;
}

/* Batch helper for bench.py's cpu_baseline / --impl reference legs: decodes captures [0,B)
 * described by offsets on `threads` POSIX threads (dynamic, one capture at a time). */
typedef struct {
    const int16_t *x; const int64_t *off; int B; const int32_t *baud; const int32_t *amp_end;
    uint8_t *out; const int64_t *out_off; afsk_oracle_result *res; int next;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    for (;;) {
        int c = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (c >= j->B) break;
        afsk_oracle_rx_decode(j->x + j->off[c], j->off[c + 1] - j->off[c], j->baud[c], j->amp_end[c],
                              j->out + j->out_off[c], 0, &j->res[c]);
    }
    return 0;
}

int afsk_oracle_rx_decode_batch(const int16_t *x, const int64_t *off, int B, const int32_t *baud,
                                const int32_t *amp_end, uint8_t *out, const int64_t *out_off,
                                afsk_oracle_result *res, int threads)
{
    batch_job j = {x, off, B, baud, amp_end, out, out_off, res, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 1; t < threads; t++) pthread_create(&th[t], 0, batch_worker, &j);
    batch_worker(&j);
    for (int t = 1; t < threads; t++) pthread_join(th[t], 0);
    return 0;
}

/* ---- Receiver.__listen :299-319, arithmetic only, over a finite recording ----
 * stream holds successive 2048-frame reads (a trailing partial chunk is ignored: pyaudio's
 * read(2048) never returns one).  Returns 1 and [start,end) if a recording was made, 0 on
 * timeout (":311-312 return []").  Finite-stream conventions (the reference would block):
 * stream exhausted before the gate opens → 0; exhausted before it closes → end = last chunk.
 */
int afsk_oracle_listen_gate(const int16_t *s, int64_t n, int amp_start, int amp_end,
                            int64_t timeout_frames, int64_t *start, int64_t *end)
{
    int64_t nchunks = n / 2048, idx = 1;                         /* :303 discard chunk 0 */
    int64_t listened = 0;
    int opened = 0;
    *start = *end = 0;
    while (listened < timeout_frames) {                          /* :304 */
        if (idx >= nchunks) return 0;
        const int16_t *fr = s + idx * 2048; idx++;               /* :305 */
        if (afsk_oracle_amplitude(fr, 2048) > amp_start) {       /* :306 */
            *start = (idx - 1) * 2048;                           /* :308 */
            opened = 1;
            break;
        }
        listened += 2048;                                        /* :310 */
    }
    if (listened >= timeout_frames || !opened) return 0;         /* :311-312 */
    *end = idx * 2048;
    while (1) {                                                  /* :313 */
        if (idx >= nchunks) break;
        const int16_t *fr = s + idx * 2048; idx++;               /* :314 */
        *end = idx * 2048;                                       /* :315 */
        if (afsk_oracle_amplitude(fr, 2048) < amp_end) break;    /* :316-318 */
    }
    return 1;
}

/* Successive Receiver.receive() calls of one receiver over one recorded stream: __listen :299-319
 * restarted at the chunk after the previous call's last read (the input stream stays open between
 * calls, :283).  out[3k..3k+2] = {recorded, start, end} of call k ("Timed out." → 0,0,0).  A call
 * that runs off the end of the recording before opening or timing out is not reported (the
 * reference would block in stream.read); a recording still open at the end keeps every full chunk. */
int64_t afsk_oracle_listen_gate_multi(const int16_t *s, int64_t n, int amp_start, int amp_end,
                                      int64_t timeout_frames, int64_t max_calls, int64_t *out)
{
    int64_t nchunks = n / 2048, idx = 0, calls = 0;
    while (calls < max_calls) {
        if (idx >= nchunks) break;
        idx++;                                                   /* :303 discard */
        int64_t listened = 0, start = 0, end = 0;
        int opened = 0, eof = 0;
        while (listened < timeout_frames) {                      /* :304 */
            if (idx >= nchunks) { eof = 1; break; }
            const int16_t *fr = s + idx * 2048; idx++;           /* :305 */
            if (afsk_oracle_amplitude(fr, 2048) > amp_start) {   /* :306 */
                start = (idx - 1) * 2048;
                opened = 1;
                break;
            }
            listened += 2048;                                    /* :310 */
        }
        if (eof) break;
        if (!opened) {                                           /* :311-312 → "Timed out." :405-407 */
            out[3 * calls] = 0; out[3 * calls + 1] = 0; out[3 * calls + 2] = 0;
            calls++;
            continue;
        }
        end = idx * 2048;
        while (1) {                                              /* :313 */
            if (idx >= nchunks) break;
            const int16_t *fr = s + idx * 2048; idx++;           /* :314 */
            end = idx * 2048;                                    /* :315 */
            if (afsk_oracle_amplitude(fr, 2048) < amp_end) break;/* :316-318 */
        }
        out[3 * calls] = 1; out[3 * calls + 1] = start; out[3 * calls + 2] = end;
        calls++;
    }
    return calls;
}

/* ---- Transmitter ---- */

/* int(baud_rate * training_time / 2) :438 is evaluated in Python (float) by the caller. */

/* number of frames __getFrames :452-469 produces BEFORE SoundOutput.__convertFrames */
int64_t afsk_oracle_tx_num_frames(int baud, int64_t ts_cycles, int64_t nbytes, const uint8_t *payload)
{
    int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (2 * AFSK_RATE + 16));
    int ns = afsk_oracle_space_tone(baud, tmp);
    int nm = afsk_oracle_mark_tone(baud, tmp);
    free(tmp);
    if (ns < 0 || nm < 0) return -1;
    int64_t total = ts_cycles * (int64_t)(nm + ns) + nm + 3 * (int64_t)ns + 4800;
    if (nm == ns) return total + nbytes * 14 * ns;
    /* unequal tones (e.g. 4800 baud): length depends on the coded bits */
    for (int64_t b = 0; b < nbytes; b++) {
        for (int half = 0; half < 2; half++) {
            int nib = half == 0 ? payload[b] >> 4 : payload[b] & 15;
            uint8_t d[4] = {(uint8_t)((nib >> 3) & 1), (uint8_t)((nib >> 2) & 1),
                            (uint8_t)((nib >> 1) & 1), (uint8_t)(nib & 1)};
            uint8_t cw[7];
            afsk_oracle_ecc_encode_nibble(d, cw);
            for (int j = 0; j < 7; j++) total += cw[j] ? nm : ns;
        }
    }
    return total;
}

/*
 * Transmitter.__getFrames :452-469 followed by SoundOutput.__convertFrames :239-244
 * (the wav payload save() :481-484 writes).  Returns frames written (len & ~1), -1 bad baud.
 */
int64_t afsk_oracle_tx_frames(const uint8_t *payload, int64_t nbytes, int baud, int64_t ts_cycles,
                              int16_t *out, int64_t cap)
{
    int16_t *space = (int16_t *)malloc(sizeof(int16_t) * (AFSK_RATE + 16));
    int16_t *mark = (int16_t *)malloc(sizeof(int16_t) * (AFSK_RATE + 16));
    int ns = afsk_oracle_space_tone(baud, space);                /* :439 */
    int nm = afsk_oracle_mark_tone(baud, mark);                  /* :440 */
    if (ns < 0 || nm < 0) { free(space); free(mark); return -1; }
    int64_t need = afsk_oracle_tx_num_frames(baud, ts_cycles, nbytes, payload);
    int16_t *fr = (int16_t *)malloc(sizeof(int16_t) * (size_t)(need + 8));
    int64_t n = 0;
#define PUSH(tone, len) do { memcpy(fr + n, tone, sizeof(int16_t) * (len)); n += (len); } while (0)
    for (int64_t i = 0; i < ts_cycles; i++) { PUSH(mark, nm); PUSH(space, ns); }   /* :457-458 */
    PUSH(mark, nm);                                              /* :460 */
    for (int i = 0; i < 3; i++) PUSH(space, ns);                 /* :461-462 */
    for (int64_t b = 0; b < nbytes; b++) {                       /* __bytesToBits :446-450: MSB first */
        uint8_t bits8[8];
        for (int j = 0; j < 8; j++) bits8[j] = (payload[b] >> (7 - j)) & 1;
        for (int half = 0; half < 2; half++) {                   /* ECC.encode :166-175 */
            uint8_t cw[7];
            afsk_oracle_ecc_encode_nibble(bits8 + 4 * half, cw);
            for (int j = 0; j < 7; j++) {                        /* :463-467 */
                if (cw[j] == 0) PUSH(space, ns); else PUSH(mark, nm);
            }
        }
    }
    memset(fr + n, 0, sizeof(int16_t) * 4800); n += 4800;        /* :468 */
#undef PUSH
    /* SoundOutput.__convertFrames :239-244: for i in range(0, len-1, 2): emit frames[i] twice */
    int64_t m = 0;
    for (int64_t i = 0; i < n - 1; i += 2) {
        if (m + 2 > cap) { m = -2; break; }
        out[m++] = fr[i];
        out[m++] = fr[i];
    }
    free(fr); free(space); free(mark);
    return m;
}
