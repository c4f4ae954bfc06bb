// Host-side codecs for the reference's on-disk formats: the LZF chunk filter h5py registers as HDF5
// filter 32000 (scripts/create_video_train_files_upsampled.py:99 `compression = 'lzf'`).  Plain C++, no device code: the
// HDF5 container logic lives in avvad/h5min.py, these two functions are its inner loops (a 5.7 MB video file is ~200
// chunks of 41 KB; byte-wise Python loops over them take seconds per file inside every DataLoader worker).
//
// Stream format (liblzf): control byte c < 32 -> c+1 literal bytes follow; otherwise a back reference of length
// (c >> 5) + 2 (a length field of 7 is extended by the next byte) at distance ((c & 31) << 8 | next byte) + 1.
// The encoder is the greedy single-probe hash matcher of that library in the configuration h5py builds it with (3-byte
// hash ((h >> (24 - hlog)) - h), hlog = 17, zero-initialised table, window 8 KiB, maximum match 264, only the last
// position of a match re-inserted), written from the format description and pinned by tests/test_h5_writer.py:
// re-compressing every chunk of the reference's shipped *.h5 files reproduces the stored bytes exactly.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {
constexpr uint32_t kMaxLit = 1u << 5;
constexpr uint32_t kMaxOff = 1u << 13;
constexpr uint32_t kMaxRef = (1u << 8) + (1u << 3);

inline uint32_t first2(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }
inline uint32_t next3(uint32_t v, const uint8_t* p) { return (v << 8) | p[2]; }
inline uint32_t slot(uint32_t h, int hlog) { return ((h >> (3 * 8 - hlog)) - h) & ((1u << hlog) - 1); }
}  // namespace

// Returns the compressed size, or 0 when the result does not fit into `cap` bytes (the caller then stores the chunk
// uncompressed, as the HDF5 filter pipeline does for an optional filter that fails).
extern "C" int64_t avvad_lzf_compress(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t cap, int hlog,
                                      uint32_t* table) {
  if (!in || !out || in_len <= 0 || cap <= 0 || hlog < 13 || hlog > 22) return 0;
  // `table` (2^hlog entries, caller-owned) lets the hash table live across calls; entries left over from an earlier
  // buffer are harmless (every candidate is verified byte by byte and must precede the cursor) and occasionally yield an
  // extra match.  h5py's build of the library does not clear its table between chunks, so carrying one table through
  // the chunks of a file in write order is what reproduces its output bit for bit; NULL = a fresh zeroed table.
  std::vector<uint32_t> own;
  if (!table) {
    own.assign((size_t)1 << hlog, 0u);
    table = own.data();
  }
  uint32_t* const htab = table;
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  uint8_t* op = out;
  uint8_t* const out_end = out + cap;
  int lit = 0;
  op++;  // room for the first literal-run header
  if (in_len < 3) {
    if (op + in_len > out_end) return 0;
  }
  uint32_t hval = in_len >= 2 ? first2(ip) : 0;
  while (ip < in_end - 2) {
    hval = next3(hval, ip);
    uint32_t* hs = &htab[slot(hval, hlog)];
    const uint8_t* ref = in + *hs;  // a stale entry may point past the cursor or the buffer: rejected by ref < ip
    *hs = (uint32_t)(ip - in);
    size_t off;
    if (ref < ip && (off = (size_t)(ip - ref - 1)) < kMaxOff && ref > in && ref[2] == ip[2] && ref[1] == ip[1] &&
        ref[0] == ip[0]) {
      uint32_t len = 2;
      uint32_t maxlen = (uint32_t)(in_end - ip) - len;
      if (maxlen > kMaxRef) maxlen = kMaxRef;
      if (op + 3 + 1 >= out_end)
        if (op - !lit + 3 + 1 >= out_end) return 0;
      op[-lit - 1] = (uint8_t)(lit - 1);  // close the literal run
      op -= !lit;                         // (or drop its header when it is empty)
      do {
        len++;
      } while (len < maxlen && ref[len] == ip[len]);
      len -= 2;  // stored length = matched bytes - 2
      ip++;
      if (len < 7) {
        *op++ = (uint8_t)((off >> 8) + (len << 5));
      } else {
        *op++ = (uint8_t)((off >> 8) + (7 << 5));
        *op++ = (uint8_t)(len - 7);
      }
      *op++ = (uint8_t)off;
      lit = 0;
      op++;  // header of the next literal run
      ip += len + 1;
      if (ip >= in_end - 2) break;
      // re-seed the hash with the position before the new cursor
      ip--;
      hval = first2(ip);
      hval = next3(hval, ip);
      htab[slot(hval, hlog)] = (uint32_t)(ip - in);
      ip++;
    } else {
      if (op >= out_end) return 0;
      lit++;
      *op++ = *ip++;
      if (lit == (int)kMaxLit) {
        op[-lit - 1] = (uint8_t)(lit - 1);
        lit = 0;
        op++;
      }
    }
  }
  if (op + 3 > out_end) return 0;
  while (ip < in_end) {
    lit++;
    *op++ = *ip++;
    if (lit == (int)kMaxLit) {
      op[-lit - 1] = (uint8_t)(lit - 1);
      lit = 0;
      op++;
    }
  }
  op[-lit - 1] = (uint8_t)(lit - 1);
  op -= !lit;
  return (int64_t)(op - out);
}

// Returns the number of bytes produced, or -1 on a malformed stream / output overflow.
extern "C" int64_t avvad_lzf_decompress(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t cap) {
  if (!in || !out || in_len < 0 || cap < 0) return -1;
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  uint8_t* op = out;
  uint8_t* const out_end = out + cap;
  while (ip < in_end) {
    uint32_t ctrl = *ip++;
    if (ctrl < 32) {
      ctrl++;
      if (op + ctrl > out_end || ip + ctrl > in_end) return -1;
      memcpy(op, ip, ctrl);
      op += ctrl;
      ip += ctrl;
    } else {
      uint32_t len = ctrl >> 5;
      if (len == 7) {
        if (ip >= in_end) return -1;
        len += *ip++;
      }
      if (ip >= in_end) return -1;
      const uint8_t* ref = op - ((ctrl & 31) << 8) - 1 - *ip++;
      len += 2;
      if (ref < out || op + len > out_end) return -1;
      for (uint32_t i = 0; i < len; ++i) op[i] = ref[i];  // may overlap: byte order matters
      op += len;
    }
  }
  return (int64_t)(op - out);
}
