// io.cc -- the input step on the host: mono 16-bit PCM WAV payloads read by
// native threads straight into the (pinned) staging buffers the copy engine
// reads from.
//
// the reference loads every utterance with scipy.io.wavfile
// inside its joblib workers (shennong/audio.py:243-286), slices segments from
// the decoded array (audio.py:520-561, utterances.py:171-176) and converts to
// int16 (audio.py:469-518).  For the common corpus format -- RIFF/WAVE, PCM,
// one channel, 16 bits -- the payload IS the int16 signal the kernels consume:
//   snb_wav_scan_batch    header walk of many files (fmt / data chunks)
//   snb_read_segments     pread of many (file, byte offset, byte count)
//                         segments into caller-owned memory
// Both run on a pool of plain threads (no Python objects, no GIL); anything
// that is not plain PCM is reported as such and goes through the Python loader.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "snb_internal.h"

namespace snb {

static uint32_t le32(const unsigned char *p) {
  return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
         (static_cast<uint32_t>(p[3]) << 24);
}
static uint32_t le16(const unsigned char *p) { return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8); }

// (data offset, samples, rate) of a mono 16-bit PCM WAV file; false for
// anything else (other encodings, channels, malformed or unreadable files)
static bool wav_layout(const char *path, int64_t *offset, int64_t *nsamples, int32_t *rate) {
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) return false;
  bool ok = false;
  unsigned char head[12], chunk[8], fmt[16];
  bool have_fmt = false;
  int64_t pos = 12;
  struct stat st;
  if (fstat(fd, &st) == 0 && pread(fd, head, 12, 0) == 12 && !std::memcmp(head, "RIFF", 4) &&
      !std::memcmp(head + 8, "WAVE", 4)) {
    while (pread(fd, chunk, 8, pos) == 8) {
      const int64_t size = le32(chunk + 4);
      pos += 8;
      if (!std::memcmp(chunk, "fmt ", 4)) {
        if (size < 16 || pread(fd, fmt, 16, pos) != 16) break;
        have_fmt = true;
      } else if (!std::memcmp(chunk, "data", 4)) {
        if (!have_fmt) break;
        if (le16(fmt) != 1 || le16(fmt + 2) != 1 || le16(fmt + 14) != 16) break;
        const int64_t left = static_cast<int64_t>(st.st_size) - pos;
        *offset = pos;
        *nsamples = (size < left ? size : left) / 2;
        *rate = static_cast<int32_t>(le32(fmt + 4));
        ok = *nsamples >= 0;
        break;
      }
      pos += size + (size & 1);
    }
  }
  close(fd);
  return ok;
}

template <typename F>
static void parallel_for(int64_t n, int32_t nthreads, F body) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = static_cast<int32_t>(n);
  if (nthreads <= 1) {
    for (int64_t i = 0; i < n; ++i) body(i);
    return;
  }
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    for (;;) {
      const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n) return;
      body(i);
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(nthreads - 1);
  for (int32_t t = 1; t < nthreads; ++t) pool.emplace_back(work);
  work();
  for (std::thread &t : pool) t.join();
}

}  // namespace snb

using namespace snb;

extern "C" int snb_wav_scan_batch(const char *const *paths, int64_t n, int64_t *data_offset, int64_t *nsamples,
                                  int32_t *rate, int32_t nthreads) {
  if (n < 0 || (n > 0 && (!paths || !data_offset || !nsamples || !rate))) return set_error(SNB_ERR_VALUE, "bad argument");
  parallel_for(n, nthreads, [&](int64_t i) {
    int64_t off = -1, ns = 0;
    int32_t r = 0;
    if (!paths[i] || !wav_layout(paths[i], &off, &ns, &r)) { off = -1; ns = 0; r = 0; }
    data_offset[i] = off; nsamples[i] = ns; rate[i] = r;
  });
  return SNB_OK;
}

extern "C" int snb_read_segments(const char *const *paths, const int64_t *offsets, const int64_t *nbytes,
                                 void *const *dst, int64_t n, int32_t nthreads, int64_t *first_failed) {
  if (n < 0 || (n > 0 && (!paths || !offsets || !nbytes || !dst))) return set_error(SNB_ERR_VALUE, "bad argument");
  std::atomic<int64_t> failed{n};
  parallel_for(n, nthreads, [&](int64_t i) {
    bool ok = false;
    const int fd = paths[i] ? open(paths[i], O_RDONLY | O_CLOEXEC) : -1;
    if (fd >= 0) {
      int64_t done = 0;
      char *out = static_cast<char *>(dst[i]);
      while (done < nbytes[i]) {
        const ssize_t got = pread(fd, out + done, static_cast<size_t>(nbytes[i] - done), offsets[i] + done);
        if (got <= 0) break;
        done += got;
      }
      ok = done == nbytes[i];
      close(fd);
    }
    if (!ok) {
      int64_t cur = failed.load();
      while (i < cur && !failed.compare_exchange_weak(cur, i)) {}
    }
  });
  const int64_t f = failed.load();
  if (first_failed) *first_failed = f < n ? f : -1;
  if (f < n) return set_error(SNB_ERR_VALUE, "%s: cannot read file, truncated data", paths[f] ? paths[f] : "(null)");
  return SNB_OK;
}
