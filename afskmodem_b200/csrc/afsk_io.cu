// afsk_io.cu — host ingest / egress of libafsk_b200.so: batched wav files <-> one pinned int16 buffer.
//
// Replaces, for n files at a time, SoundInput.loadFromFile + __convertFrames (afskmodem.py:201-205,
// 213-217: wave.open, readframes(all), pair the bytes little-endian signed WHATEVER the header says
// about channels / sample width) and SoundOutput.writeToFile (afskmodem.py:256-263: 1 channel, 2 bytes,
// 48 kHz) — 39-55 % of the reference's load() time.  Files are read by a pool of host
// threads straight into the caller's pinned buffer at their CSR offsets; finished spans are copied to
// the device as they complete, so disk/page-cache reads overlap the PCIe transfer.
//
// The RIFF walk follows CPython's wave.Wave_read.initfp / chunk.Chunk: 'RIFF' <size> 'WAVE', then
// word-aligned chunks; 'fmt ' must precede 'data'; the walk stops at 'data'.  Anything the wave module
// would reject (or any format this reader does not want to vouch for) is reported per file in
// h_status so that the Python binding can fall back to the wave module for THAT file and raise what
// the reference raises.
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "afsk_common.cuh"

namespace {

struct WavInfo {
    int64_t data_pos = 0;    // file offset of the first frame byte
    int64_t nbytes = 0;      // bytes readframes(getnframes()) returns
    int32_t status = AFSK_WAV_OK;
};

inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

bool pread_all(int fd, void *buf, size_t n, int64_t pos)
{
    uint8_t *b = static_cast<uint8_t *>(buf);
    while (n > 0) {
        const ssize_t r = pread(fd, b, n, pos);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        if (r == 0) return false;
        b += r; pos += r; n -= (size_t)r;
    }
    return true;
}

// wave.Wave_read.initfp (Lib/wave.py) restated for the PCM files the reference writes and reads
WavInfo probe_one(const char *path)
{
    WavInfo w;
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) { w.status = AFSK_WAV_E_OPEN; return w; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); w.status = AFSK_WAV_E_OPEN; return w; }
    const int64_t fsize = st.st_size;
    uint8_t h[12];
    if (fsize < 12 || !pread_all(fd, h, 12, 0) || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) {
        close(fd); w.status = AFSK_WAV_E_FORMAT; return w;
    }
    // chunk.Chunk limits every sub-chunk to the RIFF chunk's declared size
    const int64_t riff_end = std::min<int64_t>(fsize, 8 + (int64_t)rd32(h + 4));
    int64_t pos = 12;
    bool have_fmt = false;
    int framesize = 0;
    while (true) {
        uint8_t ch[8];
        if (pos + 8 > riff_end || !pread_all(fd, ch, 8, pos)) { w.status = AFSK_WAV_E_FORMAT; break; }   // fmt and/or data missing
        const int64_t csize = rd32(ch + 4);
        const int64_t body = pos + 8;
        if (memcmp(ch, "fmt ", 4) == 0) {
            uint8_t f[16];
            if (csize < 16 || body + 16 > riff_end || !pread_all(fd, f, 16, body)) { w.status = AFSK_WAV_E_FORMAT; break; }
            const int tag = rd16(f), nch = rd16(f + 2), bits = rd16(f + 14);
            const int sampwidth = (bits + 7) / 8;
            // WAVE_FORMAT_EXTENSIBLE and everything else: let the wave module decide (fallback)
            if (tag != 1 || nch == 0 || sampwidth == 0) { w.status = AFSK_WAV_E_FORMAT; break; }
            framesize = nch * sampwidth;
            have_fmt = true;
        } else if (memcmp(ch, "data", 4) == 0) {
            if (!have_fmt) { w.status = AFSK_WAV_E_FORMAT; break; }                                     // 'data chunk before fmt chunk'
            const int64_t want = (csize / framesize) * framesize;                                       // nframes * framesize
            const int64_t avail = std::max<int64_t>(0, std::min<int64_t>(riff_end, fsize) - body);      // a truncated file reads short
            w.data_pos = body;
            w.nbytes = std::min(want, avail);
            break;
        }
        pos = body + csize + (csize & 1);                                                               // chunks are word aligned
    }
    close(fd);
    return w;
}

int clamp_threads(int threads, int n)
{
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(threads, std::max(1, n)));
}

template <typename F>
void parallel_for(int n, int threads, F &&body)
{
    std::atomic<int> next{0};
    auto work = [&] {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

// ---- pinned staging ring (process-wide, grow-only) --------------------------------------------------
// afsk_wav_load with h_dst == NULL streams the files through a few pinned slots instead of pinning one
// buffer as large as the corpus: the first call of a process then costs the same as every later one
// (pinning 1.2 GB takes ~0.4 s; the ring is 32 MB by default and is kept), and the samples never need a second
// host copy.  One ring-mode load runs at a time (mutex); slot j of a call is reused by span j + R only
// after the H2D copy of span j has completed (event).
struct StageRing {
    static constexpr int kMaxSlots = 8;
    std::mutex m;
    uint8_t *slot[kMaxSlots] = {};
    size_t cap[kMaxSlots] = {};
    cudaEvent_t ev[kMaxSlots] = {};
};
StageRing g_ring;

bool ring_prepare(StageRing &r, int slots, size_t bytes)
{
    for (int j = 0; j < slots; j++) {
        if (r.cap[j] < bytes) {
            if (r.slot[j]) cudaFreeHost(r.slot[j]);
            r.slot[j] = nullptr; r.cap[j] = 0;
            if (cudaMallocHost((void **)&r.slot[j], bytes) != cudaSuccess) return false;
            r.cap[j] = bytes;
        }
        if (!r.ev[j] && cudaEventCreateWithFlags(&r.ev[j], cudaEventDisableTiming) != cudaSuccess) return false;
    }
    return true;
}

struct Piece {
    int file;
    int span;
    int64_t file_first;      // first sample of the piece inside the file's data
    int64_t stream_first;    // first sample of the piece in the concatenated stream
    int64_t count;
};

int wav_load_ring(const char *const *paths, int n, int threads, const int64_t *h_data_pos, const int64_t *h_nsamples,
                  const int64_t *h_offsets, int device, int16_t *d_dst, int64_t span_samples, cudaStream_t stream,
                  int32_t *h_status)
{
    const int64_t lo = h_offsets[0], total = h_offsets[n] - lo;
    for (int i = 0; i < n; i++) h_status[i] = AFSK_WAV_OK;
    if (total <= 0) return AFSK_OK;
    if (span_samples <= 0) {
        const char *ev = getenv("AFSK_WAV_SLOT_MB");
        // default 8 MB per slot: measured on 1024 files / 1.23 GB from /dev/shm (tools/cold_sweep.py, fresh process
        // each): 32 MB x 6 slots first call 129 ms / later 44 ms, 16 MB x 4: 72 / 30, 8 MB x 4: 55 / 29, 4 MB x 8: 57 / 29
        span_samples = (int64_t)((ev && atoi(ev) > 0) ? atoi(ev) : 8) << 19;
    }
    span_samples = std::max<int64_t>(4096, std::min<int64_t>(span_samples, (total + 4095) & ~(int64_t)4095));
    const int nspans = (int)((total + span_samples - 1) / span_samples);
    int R = 4;
    if (const char *ev = getenv("AFSK_WAV_SLOTS")) R = std::max(2, std::min(StageRing::kMaxSlots, atoi(ev)));
    R = std::min(R, std::max(nspans, 1));
    std::vector<Piece> pieces;
    std::vector<int> npieces(nspans, 0);
    for (int i = 0; i < n; i++) {
        const int64_t ns = std::min<int64_t>(h_nsamples[i], h_offsets[i + 1] - h_offsets[i]);
        int64_t done = 0;
        while (done < ns) {
            const int64_t sf = h_offsets[i] - lo + done;
            const int sp = (int)(sf / span_samples);
            const int64_t cnt = std::min<int64_t>(ns - done, (int64_t)(sp + 1) * span_samples - sf);
            pieces.push_back({i, sp, done, sf, cnt});
            npieces[sp]++;
            done += cnt;
        }
    }
    std::lock_guard<std::mutex> lock(g_ring.m);
    AfskDeviceGuard guard(device);
    if (!guard.ok) { afsk_set_error("cannot select device %d", device); return AFSK_E_CUDA; }
    if (!ring_prepare(g_ring, R, (size_t)span_samples * 2)) {
        afsk_set_error("afsk_wav_load: cannot allocate the pinned staging ring");
        return AFSK_E_CUDA;
    }
    std::vector<std::atomic<int>> remaining(nspans);
    for (int s = 0; s < nspans; s++) remaining[s].store(npieces[s]);
    std::atomic<int> released{0};        // spans whose H2D copy has completed
    std::atomic<int> next{0};
    std::atomic<bool> abort_flag{false};
    const int np = (int)pieces.size();
    auto work = [&] {
        int fd = -1, fd_file = -1;
        for (int k = next.fetch_add(1); k < np; k = next.fetch_add(1)) {
            const Piece &pc = pieces[k];
            while (pc.span >= released.load(std::memory_order_acquire) + R && !abort_flag.load())
                std::this_thread::sleep_for(std::chrono::microseconds(20));
            if (!abort_flag.load()) {
                if (fd_file != pc.file) {
                    if (fd >= 0) close(fd);
                    fd = open(paths[pc.file], O_RDONLY | O_CLOEXEC);
                    fd_file = pc.file;
                }
                uint8_t *dst = g_ring.slot[pc.span % R] + (size_t)(pc.stream_first - (int64_t)pc.span * span_samples) * 2;
                if (fd < 0 || !pread_all(fd, dst, (size_t)pc.count * 2, h_data_pos[pc.file] + pc.file_first * 2))
                    h_status[pc.file] = AFSK_WAV_E_OPEN;
            }
            remaining[pc.span].fetch_sub(1, std::memory_order_release);
        }
        if (fd >= 0) close(fd);
    };
    if (threads <= 0) threads = std::max(1, (int)std::thread::hardware_concurrency() - 2);
    threads = clamp_threads(threads, np);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    int rc = AFSK_OK;
    int issued = 0;
    auto poll_released = [&] {
        int r = released.load();
        while (r < issued && cudaEventQuery(g_ring.ev[r % R]) == cudaSuccess) r++;
        released.store(r, std::memory_order_release);
    };
    for (int s = 0; s < nspans && rc == AFSK_OK; s++) {
        while (remaining[s].load(std::memory_order_acquire) > 0) {
            poll_released();
            std::this_thread::sleep_for(std::chrono::microseconds(20));       // the readers need the cores
        }
        const int64_t a = (int64_t)s * span_samples, b = std::min<int64_t>(total, a + span_samples);
        if (cudaMemcpyAsync(d_dst + lo + a, g_ring.slot[s % R], (size_t)(b - a) * 2, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
            cudaEventRecord(g_ring.ev[s % R], stream) != cudaSuccess) {
            afsk_set_error("afsk_wav_load: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = AFSK_E_CUDA;
            abort_flag.store(true);
            break;
        }
        issued = s + 1;
        poll_released();
    }
    for (auto &t : pool) t.join();
    // the slots belong to the process-wide ring: they may be refilled as soon as this call returns
    for (int s = std::max(0, issued - R); s < issued; s++) cudaEventSynchronize(g_ring.ev[s % R]);
    return rc;
}

}  // namespace

extern "C" {

int afsk_wav_probe(const char *const *paths, int n, int threads, int64_t *h_nsamples, int64_t *h_data_pos, int32_t *h_status)
{
    if (n < 0 || (n > 0 && (!paths || !h_nsamples || !h_data_pos || !h_status))) return AFSK_E_ARG;
    parallel_for(n, clamp_threads(threads, n), [&](int i) {
        const WavInfo w = probe_one(paths[i]);
        h_status[i] = w.status;
        h_data_pos[i] = w.data_pos;
        h_nsamples[i] = w.status == AFSK_WAV_OK ? w.nbytes / 2 : 0;      // __convertFrames :203 pairs bytes, odd tail dropped
    });
    return AFSK_OK;
}

int afsk_wav_load(const char *const *paths, int n, int threads, const int64_t *h_data_pos, const int64_t *h_nsamples,
                  const int64_t *h_offsets, int16_t *h_dst, int device, int16_t *d_dst, int64_t span_samples, void *stream,
                  int32_t *h_status)
{
    if (n < 0 || (n > 0 && (!paths || !h_data_pos || !h_nsamples || !h_offsets || !h_status))) return AFSK_E_ARG;
    if (n > 0 && !h_dst && !d_dst) { afsk_set_error("afsk_wav_load: neither a host nor a device destination"); return AFSK_E_ARG; }
    if (n == 0) return AFSK_OK;
    if (!h_dst) return wav_load_ring(paths, n, threads, h_data_pos, h_nsamples, h_offsets, device, d_dst, span_samples, (cudaStream_t)stream, h_status);
    // automatic thread count: leave two cores to the copy-issuing thread and the driver when the H2D is
    // overlapped (measured on a 16-core host, 1.23 GB from /dev/shm: 12 readers 31.2 ms, 16 readers 32.4 ms,
    // 24 readers 36.6 ms; without the overlapped copy 16 readers fill the buffer in 22.2 ms)
    if (threads <= 0 && d_dst) threads = std::max(1, (int)std::thread::hardware_concurrency() - 2);
    threads = clamp_threads(threads, n);
    if (span_samples <= 0) span_samples = (int64_t)32 << 20;            // 64 MB per H2D copy
    // spans of consecutive files, each copied to the device as soon as its last file is in memory
    std::vector<int> span_first;
    for (int i = 0; i < n;) {
        span_first.push_back(i);
        const int64_t base = h_offsets[i];
        int j = i + 1;
        while (j < n && h_offsets[j + 1] - base <= span_samples) j++;
        i = j;
    }
    span_first.push_back(n);
    const int nspans = (int)span_first.size() - 1;
    std::vector<int> file_span(n);
    std::vector<std::atomic<int>> remaining(nspans);
    for (int s = 0; s < nspans; s++) {
        remaining[s].store(span_first[s + 1] - span_first[s]);
        for (int i = span_first[s]; i < span_first[s + 1]; i++) file_span[i] = s;
    }
    std::atomic<int> next{0};
    auto work = [&] {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
            int32_t st = AFSK_WAV_OK;
            const int64_t ns = std::min<int64_t>(h_nsamples[i], h_offsets[i + 1] - h_offsets[i]);
            if (ns > 0) {
                const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
                if (fd < 0 || !pread_all(fd, h_dst + h_offsets[i], (size_t)ns * 2, h_data_pos[i])) st = AFSK_WAV_E_OPEN;
                if (fd >= 0) close(fd);
            }
            h_status[i] = st;
            remaining[file_span[i]].fetch_sub(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    int rc = AFSK_OK;
    if (d_dst) {
        AfskDeviceGuard guard(device);
        if (!guard.ok) rc = AFSK_E_CUDA;
        for (int s = 0; s < nspans && rc == AFSK_OK; s++) {
            while (remaining[s].load(std::memory_order_acquire) > 0)
                std::this_thread::sleep_for(std::chrono::microseconds(20));   // the readers need the cores
            const int64_t a = h_offsets[span_first[s]], b = h_offsets[span_first[s + 1]];
            if (b > a && cudaMemcpyAsync(d_dst + a, h_dst + a, (size_t)(b - a) * 2, cudaMemcpyHostToDevice,
                                         (cudaStream_t)stream) != cudaSuccess) {
                afsk_set_error("afsk_wav_load: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = AFSK_E_CUDA;
            }
        }
    }
    for (auto &t : pool) t.join();
    return rc;
}

int afsk_wav_save(const char *const *paths, int n, int threads, const int16_t *h_src, const int64_t *h_start,
                  const int64_t *h_len, int32_t *h_status)
{
    if (n < 0 || (n > 0 && (!paths || !h_src || !h_start || !h_len || !h_status))) return AFSK_E_ARG;
    parallel_for(n, clamp_threads(threads, n), [&](int i) {
        // wave.Wave_write header for setnchannels(1) / setsampwidth(2) / setframerate(48000) (:258-261)
        const uint64_t bytes = (uint64_t)h_len[i] * 2;
        h_status[i] = AFSK_WAV_OK;
        if (bytes + 36 > 0xFFFFFFFFull) { h_status[i] = AFSK_WAV_E_FORMAT; return; }    // struct.pack('<L') would raise
        uint8_t h[44];
        auto w32 = [&](int o, uint32_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; h[o + 2] = (v >> 16) & 255; h[o + 3] = v >> 24; };
        auto w16 = [&](int o, uint16_t v) { h[o] = v & 255; h[o + 1] = v >> 8; };
        memcpy(h, "RIFF", 4); w32(4, (uint32_t)(36 + bytes)); memcpy(h + 8, "WAVEfmt ", 8); w32(16, 16);
        w16(20, 1); w16(22, 1); w32(24, AFSK_RATE); w32(28, AFSK_RATE * 2); w16(32, 2); w16(34, 16);
        memcpy(h + 36, "data", 4); w32(40, (uint32_t)bytes);
        const int fd = open(paths[i], O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
        if (fd < 0) { h_status[i] = AFSK_WAV_E_OPEN; return; }
        bool ok = write(fd, h, 44) == 44;
        const uint8_t *p = reinterpret_cast<const uint8_t *>(h_src + h_start[i]);
        uint64_t left = bytes;
        while (ok && left > 0) {
            const ssize_t r = write(fd, p, (size_t)std::min<uint64_t>(left, 1u << 30));
            if (r < 0) { if (errno == EINTR) continue; ok = false; break; }
            p += r; left -= (uint64_t)r;
        }
        if (close(fd) != 0) ok = false;
        if (!ok) h_status[i] = AFSK_WAV_E_OPEN;
    });
    return AFSK_OK;
}

}  // extern "C"
