/*
 * afsk_b200.h — C ABI of libafsk_b200.so, the B200 (sm_100a) batch AFSK modem core.
 *
 * The reference (lavajuno/afskmodem, one pure-Python file) has no FFI/plugin interface; its
 * boundary is the Python class API (afskmodem.py:275-276 Receiver, :420 load, :437 Transmitter,
 * :481 save).  This header is the C surface a binding for that path needs: plain pointers and
 * sizes, no torch types.  Each entry point names the reference function(s) it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (AFSK_E_*); afsk_last_error() gives text
 *     (thread-local).  Nothing throws.  There is NO CPU fallback: without a CUDA device every
 *     compute entry point returns AFSK_E_CUDA.
 *   - device pointers are caller-owned; `stream` is a cudaStream_t passed as void* (NULL =
 *     default stream).  Calls taking a stream are asynchronous on it.
 *   - sample buffers are int16 little-endian, 48 kHz mono, captures concatenated; offsets are in
 *     samples (CSR: capture c = [off[c], off[c+1]) ).  The device sample pointer must be 16-byte
 *     aligned and the allocation must extend to the next 16-byte boundary past off[B].
 *   - one process/thread per GPU; a plan belongs to the device it was created on.
 */
#ifndef AFSK_B200_H
#define AFSK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFSK_ABI_VERSION 2

/* return codes */
#define AFSK_OK 0
#define AFSK_E_ARG (-1)     /* bad argument (null, misaligned, negative size ...) */
#define AFSK_E_CUDA (-2)    /* CUDA runtime error / no device                      */
#define AFSK_E_BAUD (-3)    /* Exception("Invalid baud rate.")  afskmodem.py:69-70,81-82 */
#define AFSK_E_UNSUPPORTED (-4)

/* per-capture decode status (AfskRxResult.status) */
#define AFSK_ST_OK 0            /* >= 1 coded bit decoded                                      */
#define AFSK_ST_NO_CLOCK 1      /* len(frames) < 4096: "Failed to recover clock" :323-325      */
#define AFSK_ST_NO_DATA 2       /* clock found but 0 bits: "No data." :422-424                 */
#define AFSK_ST_EXC_WAVELEN (-1)/* Exception("Comparing two waveforms of different lengths.")  */
#define AFSK_ST_EXC_INDEX (-2)  /* IndexError at scan_diffs[0] :332 (2*bit_frames >= 4096)      */
#define AFSK_ST_EXC_BAUD (-3)   /* constructor would raise "Invalid baud rate."                */

/* One per capture; the four integers are the reference's debug-log stage values
 * (afskmodem.py:338, :368, :380, :427). */
typedef struct AfskRxResult {
    int32_t status;
    int32_t clock;       /* -1 when no clock */
    int64_t train_end;   /* -1 when no clock */
    int64_t nbits;
    int64_t nbytes;
} AfskRxResult;

typedef struct AfskRxPlan AfskRxPlan;
typedef struct AfskTxPlan AfskTxPlan;

/* ---------------------------------------------------------------- library / device ---- */
int afsk_abi_version(void);
const char *afsk_last_error(void);
int afsk_device_count(int *count);
/* SM count, global memory bytes, compute capability major*10+minor of `device` */
int afsk_device_info(int device, int *sm_count, size_t *mem_bytes, int *cc);

/* minimal memory / stream helpers so a host binding needs no other CUDA wrapper */
int afsk_malloc(int device, size_t bytes, void **dptr);
int afsk_free(int device, void *dptr);
int afsk_host_alloc(size_t bytes, void **hptr);           /* pinned */
int afsk_host_free(void *hptr);
int afsk_memcpy_h2d(int device, void *dst, const void *src, size_t bytes, void *stream);
int afsk_memcpy_d2h(int device, void *dst, const void *src, size_t bytes, void *stream);
int afsk_memset(int device, void *dst, int value, size_t bytes, void *stream);
int afsk_stream_create(int device, void **stream);
int afsk_stream_destroy(int device, void *stream);
int afsk_stream_sync(int device, void *stream);
/* everything enqueued on `waiter` after this call runs after everything enqueued on `signaller`
 * before it (event record + stream wait): lets a copy stream feed a compute stream chunk by chunk */
int afsk_stream_wait_stream(int device, void *waiter, void *signaller);

/* ---------------------------------------------------------------- tone tables ---------- */
/* Waveforms.getSpaceTone / getMarkTone lengths (afskmodem.py:68-85) and Receiver.__bit_frames
 * (:277).  Returns AFSK_E_BAUD where the reference constructor raises. */
int afsk_tone_lengths(int baud, int *bit_frames, int *mark_len, int *space_len);

/* ---------------------------------------------------------------- receiver ------------- */
/*
 * Replaces Receiver.load's compute (afskmodem.py:420-430): __decodeBits :354-381
 * (__recoverClockIndex :322-339, __decodeBit :342-351 with __amplify :287-296 and
 * Waveforms.getDiff/getAmplitude :94-107, __scanTraining :386-390), ECC.decode :154-163 and
 * __bitsToBytes :393-399, for B independent captures.
 *
 * h_offsets[B+1] (samples), h_baud[B], h_amp_end[B] are HOST arrays describing the batch
 * (amp_end = Receiver's amp_end_threshold; amp_start is unused on the file path, :375 vs :306).
 * A plan holds the device-side descriptors and scratch (bit planes, clock indices).
 */
int afsk_rx_plan_create(int device, int B, const int64_t *h_offsets, const int32_t *h_baud,
                        const int32_t *h_amp_end, AfskRxPlan **plan);
/* same, for captures given as arbitrary (possibly non-adjacent) [start, start + len) ranges of one
 * sample buffer — e.g. the recordings afsk_rx_gate_multi cuts out of a long stream */
int afsk_rx_plan_create_ranges(int device, int B, const int64_t *h_start, const int64_t *h_len,
                               const int32_t *h_baud, const int32_t *h_amp_end, AfskRxPlan **plan);
/* Re-targets an existing plan at another batch layout (a drop-in Receiver.load sees a different file
 * length on every call): same arguments as afsk_rx_plan_create_ranges, or h_len == NULL with CSR
 * h_start[B+1].  The plan's device arena is reused when large enough (it only grows), so the call costs
 * one host-to-device copy of the descriptors and no allocation.  Output offsets / launch counts queried
 * earlier are invalidated. */
int afsk_rx_plan_reset(AfskRxPlan *plan, int B, const int64_t *h_start, const int64_t *h_len,
                       const int32_t *h_baud, const int32_t *h_amp_end);
int afsk_rx_plan_destroy(AfskRxPlan *plan);
/* tuning / test switches of a plan; results never depend on them */
#define AFSK_OPT_FRAME_KERNEL 1 /* framing kernel: 0 automatic (by capture length), 1 k_frame_warp (one warp per
                                   capture), 2 k_frame<128,4>, 3 k_frame<512,8>, 4 k_frame<512,8> with 64-bit window
                                   indices (automatic from 2^30 windows per capture).  A variant that cannot
                                   represent the batch is replaced by the automatic choice. */
#define AFSK_OPT_L2_HINT 2      /* L2 evict-first hint on the demodulator's bulk copies: -1 per-kernel default, 0, 1 */
#define AFSK_OPT_FUSED 3        /* schedule of a decode.  0: three kernels (clock, demodulate, frame).  1: auxiliary warps
                                   of the streaming kernel recover the clocks (no k_clock launch).  2: they also frame
                                   every capture as soon as its last tile has retired (one launch per bit length).
                                   -1 automatic = 0 (neither fused level beat the three kernels on a B200, see
                                   profiles/r2_tuning_log.md; both stay available and parity-tested).  Bit lengths over
                                   185 frames (below 260 baud) keep the three kernels; captures of more than 2^18 bit
                                   windows keep the separate framing kernel. */
#define AFSK_OPT_CLOCK_KERNEL 4 /* clock recovery kernel of the three-kernel schedule: 0 (default) automatic = k_clock_q
                                   (32 candidates per thread from registers, no prefix array) for bit lengths of 8..24
                                   frames (6000..2000 baud) and k_clock for the rest; 1 k_clock everywhere (two sweeps
                                   over candidate distances kept in registers); 2 k_clock2 (one sweep with an exact
                                   multiply-shift floor, half the shared memory; bit lengths up to 185 frames, else k_clock) */
#define AFSK_OPT_GROUP_STREAMS 5 /* 1 (default): the demodulator launches of a mixed-baud batch (one per bit length) run on
                                   streams of their own between the clock and framing kernels, so that one group's CTAs
                                   fill the SMs as the previous group's drain; 0: one after the other on the caller's stream */
int afsk_rx_plan_set_option(AfskRxPlan *plan, int option, int value);
/* capacity offsets (bytes, B+1 entries, host memory owned by the plan) of the decoded output */
int afsk_rx_plan_out_offsets(const AfskRxPlan *plan, const int64_t **h_out_off);
/* number of kernel launches one afsk_rx_decode issues for this plan */
int afsk_rx_plan_launches(const AfskRxPlan *plan, int *launches);
/*
 * d_samples: device int16[h_offsets[B]] ; d_out: device bytes[h_out_off[B]] ; d_res: device
 * AfskRxResult[B].  Capture c's payload is d_out[h_out_off[c] .. + d_res[c].nbytes).
 */
int afsk_rx_decode(AfskRxPlan *plan, const int16_t *d_samples, uint8_t *d_out, AfskRxResult *d_res,
                   void *stream);
/*
 * Optional in-line timing of the dominant kernel (k_demod) for roofline reporting: when enabled,
 * every afsk_rx_decode brackets each k_demod launch with CUDA events on the caller's stream.
 * afsk_rx_plan_demod_time synchronizes those events, returns their summed duration and launch
 * count since the last query, and clears them.
 */
int afsk_rx_plan_set_timing(AfskRxPlan *plan, int enable);
int afsk_rx_plan_demod_time(AfskRxPlan *plan, float *ms_total, int *launches);
/* raw decision / quiet flags of capture c after a decode, for diagnostics and stage-level parity
 * tests: d_planes points at interleaved uint32 pairs {bits, quiet}; window k of the capture is
 * bit (k & 31) of pair k >> 5 (LSB first). */
int afsk_rx_plan_planes(const AfskRxPlan *plan, int capture, const uint32_t **d_planes, int64_t *max_windows);
/*
 * Host-buffer convenience (what a drop-in Receiver.load calls): H2D, decode, D2H on `device`.
 * h_out_off[B+1] are capacity offsets into h_out (use afsk_rx_out_capacity per capture).
 * Calls on one device are serialised; the plan and buffers of the previous call are reused.
 */
int afsk_rx_decode_host(int device, const int16_t *h_samples, const int64_t *h_offsets, int B,
                        const int32_t *h_baud, const int32_t *h_amp_end, uint8_t *h_out,
                        const int64_t *h_out_off, AfskRxResult *h_res);
/* afsk_rx_decode_host keeps a plan and grow-only device / pinned buffers per device between calls;
 * this releases them (also done at process exit by the driver) */
int afsk_rx_host_release(int device);
/* upper bound on decoded bytes of a capture of n samples at `baud` */
int64_t afsk_rx_out_capacity(int64_t n_samples, int baud);

/*
 * Receiver.__listen gate arithmetic (afskmodem.py:299-319) over S recorded streams: chunk 0 is
 * discarded, the first 2048-frame chunk with floor(sum|x|/2048) > amp_start opens the recording,
 * it extends through the first chunk with amplitude < amp_end.  h_offsets[S+1] in samples.
 * d_range receives int64 {recorded(0/1), start, end} per stream.  Finite-stream conventions:
 * exhausted before opening → recorded = 0; exhausted before closing → end = last full chunk.
 */
int afsk_rx_gate(int device, const int16_t *d_samples, const int64_t *h_offsets, int S, int amp_start,
                 int amp_end, int64_t timeout_frames, int64_t *d_range, void *stream);

/*
 * Successive Receiver.receive(timeout) calls of ONE receiver over each recorded stream (the
 * reference keeps its input stream open between calls, afskmodem.py:283, so call k+1 starts at the
 * chunk after call k's last read): per call one int64 triple {recorded(0 = "Timed out."), start, end}
 * in d_ranges[s * max_calls * 3 ...], calls made per stream in d_counts[s].  A call that reaches
 * the end of the recording before opening or timing out is not reported.
 */
int afsk_rx_gate_multi(int device, const int16_t *d_samples, const int64_t *h_offsets, int S, int amp_start,
                       int amp_end, int64_t timeout_frames, int max_calls, int64_t *d_ranges,
                       int32_t *d_counts, void *stream);

/* ---------------------------------------------------------------- transmitter ---------- */
/* frames Transmitter.save writes for a payload of n bytes (afskmodem.py:452-469 then :239-244).
 * Needs the payload only when mark/space tone lengths differ (e.g. 4800 baud); pass NULL else. */
int64_t afsk_tx_num_samples(int baud, int64_t ts_cycles, int64_t payload_bytes, const uint8_t *payload);
/*
 * Replaces Transmitter.__getFrames (:452-469: __bytesToBits :446-450, ECC.encode :166-175,
 * training sequence, terminator, 4800-frame tail) + SoundOutput.__convertFrames (:239-244) for
 * B payloads.  h_pay_off[B+1] bytes; h_baud[B]; h_ts_cycles[B] = int(baud*training_time/2) (:438).
 */
int afsk_tx_plan_create(int device, int B, const int64_t *h_pay_off, const int32_t *h_baud,
                        const int64_t *h_ts_cycles, const uint8_t *h_payload, AfskTxPlan **plan);
int afsk_tx_plan_destroy(AfskTxPlan *plan);
/* sample offsets (B+1, host, owned by the plan; each capture starts on a multiple of 8 samples) */
int afsk_tx_plan_out_offsets(const AfskTxPlan *plan, const int64_t **h_out_off, const int64_t **h_out_len);
int afsk_tx_synth(AfskTxPlan *plan, const uint8_t *d_payload, int16_t *d_out, void *stream);
int afsk_tx_synth_host(int device, const uint8_t *h_payload, const int64_t *h_pay_off, int B,
                       const int32_t *h_baud, const int64_t *h_ts_cycles, int16_t *h_out,
                       const int64_t *h_out_off);

/* ---------------------------------------------------------------- wav files ------------ */
/*
 * Host ingest / egress for batches of files (one pool of host threads; threads <= 0: all cores).
 * Replaces SoundInput.loadFromFile + __convertFrames (afskmodem.py:201-205, 213-217: every frame byte
 * of the data chunk, paired little-endian signed whatever the header says about channels or sample
 * width) and SoundOutput.writeToFile (afskmodem.py:256-263: 1 channel / 2 bytes / 48000 Hz).
 * Per-file status: AFSK_WAV_OK, or a reason to hand the file to CPython's wave module instead
 * (which then raises exactly what the reference raises).
 */
#define AFSK_WAV_OK 0
#define AFSK_WAV_E_OPEN 1      /* cannot open / read / write the file                     */
#define AFSK_WAV_E_FORMAT 2    /* not plain RIFF/WAVE PCM with fmt before data            */
/* samples (= data bytes / 2) and file position of the first frame byte of each file */
int afsk_wav_probe(const char *const *paths, int n, int threads, int64_t *h_nsamples, int64_t *h_data_pos,
                   int32_t *h_status);
/* reads file i's samples into h_dst[h_offsets[i] ...] (pinned memory recommended).  With d_dst != NULL,
 * spans of about span_samples consecutive samples are copied to d_dst (same offsets) on `stream` as
 * soon as their files are in memory, so that reading overlaps the PCIe transfer.
 * h_dst == NULL (d_dst required): the files stream through a process-wide ring of pinned staging slots
 * (span_samples each, 8 MB by default, four slots) straight to the device; no host copy of the corpus is kept, the
 * first call of a process costs what later calls cost, and the call returns once every copy has landed. */
int afsk_wav_load(const char *const *paths, int n, int threads, const int64_t *h_data_pos, const int64_t *h_nsamples,
                  const int64_t *h_offsets, int16_t *h_dst, int device, int16_t *d_dst, int64_t span_samples,
                  void *stream, int32_t *h_status);
/* writes h_src[h_start[i] .. + h_len[i]) as file i, byte-identical to the reference's writeToFile */
int afsk_wav_save(const char *const *paths, int n, int threads, const int16_t *h_src, const int64_t *h_start,
                  const int64_t *h_len, int32_t *h_status);

#ifdef __cplusplus
}
#endif
#endif /* AFSK_B200_H */
