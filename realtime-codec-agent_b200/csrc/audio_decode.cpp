// Host-side FLAC decoder of the corpus ingest path (libri-light, one of the corpora the reference encodes, ships as
// .flac; the reference loads it through librosa -> soundfile -> libFLAC, prep_channel_map.py:7-8 / the codec_bpe CLI).
// No audio library is installed in this image, so the container is decoded here from the published format
// (RFC 9639): STREAMINFO, frame headers with CRC-8, CONSTANT / VERBATIM / FIXED / LPC subframes, Rice and Rice2
// residuals with escape partitions, wasted bits, left/side - right/side - mid/side stereo, frame CRC-16.
// Output is interleaved PCM ready for mc_op_pcm_to_f32: int16 for streams of <= 16 bits, int32 (left-justified)
// above.  Pure C++ (no CUDA); exported from the same shared library as the engine.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/magicodec_b200.h"

namespace {

struct BitReader {
  const uint8_t* p;
  int64_t n;      // bytes
  int64_t pos;    // bit position
  bool fail = false;
  BitReader(const uint8_t* d, int64_t bytes, int64_t byte_pos) : p(d), n(bytes), pos(byte_pos * 8) {}
  inline uint32_t bit() {
    if ((pos >> 3) >= n) { fail = true; return 0; }
    const uint32_t b = (p[pos >> 3] >> (7 - (pos & 7))) & 1u;
    ++pos;
    return b;
  }
  inline uint64_t bits(int k) {   // k <= 57
    if (k == 0) return 0;
    if (((pos + k + 7) >> 3) > n) { fail = true; pos += k; return 0; }
    uint64_t v = 0;
    int need = k;
    while (need > 0) {
      const int off = static_cast<int>(pos & 7);
      const int take = need < 8 - off ? need : 8 - off;
      const uint32_t byte = p[pos >> 3];
      v = (v << take) | ((byte >> (8 - off - take)) & ((1u << take) - 1u));
      pos += take;
      need -= take;
    }
    return v;
  }
  inline int64_t sbits(int k) {
    if (k == 0) return 0;
    const uint64_t v = bits(k);
    const uint64_t m = 1ull << (k - 1);
    return static_cast<int64_t>((v ^ m)) - static_cast<int64_t>(m);
  }
  inline uint32_t unary() {   // zeros before the first one
    uint32_t q = 0;
    // fast path: scan whole bytes
    while (true) {
      if ((pos >> 3) >= n) { fail = true; return q; }
      const int off = static_cast<int>(pos & 7);
      const uint32_t rest = static_cast<uint32_t>(p[pos >> 3] << off) & 0xFFu;   // remaining bits of this byte, left aligned
      if (rest == 0) { q += 8 - off; pos += 8 - off; continue; }
      const int lz = __builtin_clz(rest) - 24;
      q += lz;
      pos += lz + 1;
      return q;
    }
  }
  inline void align() { pos = (pos + 7) & ~int64_t(7); }
  inline int64_t byte_pos() const { return pos >> 3; }
};

uint8_t crc8(const uint8_t* d, int64_t n) {
  uint8_t c = 0;
  for (int64_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int b = 0; b < 8; ++b) c = (c & 0x80) ? static_cast<uint8_t>((c << 1) ^ 0x07) : static_cast<uint8_t>(c << 1);
  }
  return c;
}

uint16_t g_crc16_table[256];
bool g_crc16_ready = false;
uint16_t crc16(const uint8_t* d, int64_t n) {
  if (!g_crc16_ready) {
    for (int i = 0; i < 256; ++i) {
      uint16_t c = static_cast<uint16_t>(i << 8);
      for (int b = 0; b < 8; ++b) c = (c & 0x8000) ? static_cast<uint16_t>((c << 1) ^ 0x8005) : static_cast<uint16_t>(c << 1);
      g_crc16_table[i] = c;
    }
    g_crc16_ready = true;
  }
  uint16_t c = 0;
  for (int64_t i = 0; i < n; ++i) c = static_cast<uint16_t>((c << 8) ^ g_crc16_table[(c >> 8) ^ d[i]]);
  return c;
}

struct StreamInfo {
  int32_t sample_rate = 0, channels = 0, bits = 0;
  int64_t total = 0;
  int64_t first_frame = 0;   // byte offset of the first audio frame
  int32_t max_block = 0;
};

int parse_streaminfo(const uint8_t* d, int64_t n, StreamInfo* si) {
  int64_t pos = 0;
  // an ID3v2 tag may precede the marker
  if (n >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3') {
    const int64_t sz = (int64_t(d[6] & 0x7F) << 21) | (int64_t(d[7] & 0x7F) << 14) | (int64_t(d[8] & 0x7F) << 7) | int64_t(d[9] & 0x7F);
    pos = 10 + sz;
  }
  if (pos + 4 > n || memcmp(d + pos, "fLaC", 4) != 0) return MC_ERR_ARG;
  pos += 4;
  bool have = false;
  while (true) {
    if (pos + 4 > n) return MC_ERR_ARG;
    const bool last = (d[pos] & 0x80) != 0;
    const int type = d[pos] & 0x7F;
    const int64_t len = (int64_t(d[pos + 1]) << 16) | (int64_t(d[pos + 2]) << 8) | int64_t(d[pos + 3]);
    pos += 4;
    if (pos + len > n) return MC_ERR_ARG;
    if (type == 0) {
      if (len < 34) return MC_ERR_ARG;
      const uint8_t* s = d + pos;
      si->max_block = (s[2] << 8) | s[3];
      si->sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
      si->channels = ((s[12] >> 1) & 7) + 1;
      si->bits = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
      si->total = (int64_t(s[13] & 0x0F) << 32) | (int64_t(s[14]) << 24) | (int64_t(s[15]) << 16) | (int64_t(s[16]) << 8) | int64_t(s[17]);
      have = true;
    }
    pos += len;
    if (last) break;
  }
  if (!have || si->sample_rate <= 0 || si->bits < 4 || si->bits > 32) return MC_ERR_ARG;
  si->first_frame = pos;
  return MC_OK;
}

// residual of one subframe into out[order .. blocksize)
bool read_residual(BitReader& br, int32_t* out, int blocksize, int order) {
  const int method = static_cast<int>(br.bits(2));
  if (method > 1) return false;
  const int pbits = method == 0 ? 4 : 5;
  const int escape = method == 0 ? 15 : 31;
  const int porder = static_cast<int>(br.bits(4));
  const int parts = 1 << porder;
  if ((blocksize >> porder) << porder != blocksize && porder > 0) return false;
  int idx = order;
  for (int p = 0; p < parts; ++p) {
    int count = (porder == 0) ? blocksize - order : ((blocksize >> porder) - (p == 0 ? order : 0));
    if (count < 0 || idx + count > blocksize) return false;
    const int k = static_cast<int>(br.bits(pbits));
    if (k == escape) {
      const int raw = static_cast<int>(br.bits(5));
      for (int i = 0; i < count; ++i) out[idx++] = static_cast<int32_t>(br.sbits(raw));
    } else {
      for (int i = 0; i < count; ++i) {
        const uint32_t q = br.unary();
        const uint32_t u = (q << k) | static_cast<uint32_t>(br.bits(k));
        out[idx++] = static_cast<int32_t>(u >> 1) ^ -static_cast<int32_t>(u & 1);
      }
    }
    if (br.fail) return false;
  }
  return idx == blocksize;
}

bool read_subframe(BitReader& br, int32_t* out, int blocksize, int bps) {
  if (br.bit() != 0) return false;
  const int type = static_cast<int>(br.bits(6));
  int wasted = 0;
  if (br.bit()) wasted = static_cast<int>(br.unary()) + 1;
  bps -= wasted;
  if (bps < 1 || br.fail) return false;
  if (type == 0) {                               // CONSTANT
    const int32_t v = static_cast<int32_t>(br.sbits(bps));
    for (int i = 0; i < blocksize; ++i) out[i] = v;
  } else if (type == 1) {                        // VERBATIM
    for (int i = 0; i < blocksize; ++i) out[i] = static_cast<int32_t>(br.sbits(bps));
  } else if (type >= 8 && type <= 12) {          // FIXED, order type - 8
    const int order = type - 8;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) out[i] = static_cast<int32_t>(br.sbits(bps));
    if (!read_residual(br, out, blocksize, order)) return false;
    for (int i = order; i < blocksize; ++i) {
      int64_t pred = 0;
      switch (order) {
        case 1: pred = out[i - 1]; break;
        case 2: pred = 2 * int64_t(out[i - 1]) - out[i - 2]; break;
        case 3: pred = 3 * int64_t(out[i - 1]) - 3 * int64_t(out[i - 2]) + out[i - 3]; break;
        case 4: pred = 4 * int64_t(out[i - 1]) - 6 * int64_t(out[i - 2]) + 4 * int64_t(out[i - 3]) - out[i - 4]; break;
        default: break;
      }
      out[i] = static_cast<int32_t>(pred + out[i]);
    }
  } else if (type >= 32) {                       // LPC, order (type & 31) + 1
    const int order = (type & 31) + 1;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) out[i] = static_cast<int32_t>(br.sbits(bps));
    const int prec = static_cast<int>(br.bits(4)) + 1;
    if (prec == 16) return false;
    const int shift = static_cast<int>(br.sbits(5));
    if (shift < 0) return false;
    int32_t coef[32];
    for (int j = 0; j < order; ++j) coef[j] = static_cast<int32_t>(br.sbits(prec));
    if (!read_residual(br, out, blocksize, order)) return false;
    for (int i = order; i < blocksize; ++i) {
      int64_t acc = 0;
      for (int j = 0; j < order; ++j) acc += int64_t(coef[j]) * out[i - 1 - j];
      out[i] = static_cast<int32_t>((acc >> shift) + out[i]);
    }
  } else {
    return false;                                // reserved subframe type
  }
  if (wasted)
    for (int i = 0; i < blocksize; ++i) out[i] = static_cast<int32_t>(static_cast<uint32_t>(out[i]) << wasted);
  return !br.fail;
}

}  // namespace

extern "C" {

int mc_flac_info(const uint8_t* data, int64_t n, int32_t* sample_rate, int32_t* channels, int32_t* bits, int64_t* total_samples) {
  if (!data || n < 42) return MC_ERR_ARG;
  StreamInfo si;
  const int rc = parse_streaminfo(data, n, &si);
  if (rc != MC_OK) return rc;
  if (sample_rate) *sample_rate = si.sample_rate;
  if (channels) *channels = si.channels;
  if (bits) *bits = si.bits;
  if (total_samples) *total_samples = si.total;
  return MC_OK;
}

/* out: interleaved [frames][channels]; int16 when the stream has <= 16 bits per sample, else int32 left-justified
 * (sample << (32 - bits)).  capacity_frames = room in `out`; *decoded = frames written.  Frames whose CRC-16 does not
 * match, or a stream that ends early, return MC_ERR_ARG (a truncated corpus file must not pass silently). */
int mc_flac_decode(const uint8_t* data, int64_t n, void* out, int64_t capacity_frames, int64_t* decoded) {
  if (!data || !out || !decoded) return MC_ERR_ARG;
  *decoded = 0;
  StreamInfo si;
  int rc = parse_streaminfo(data, n, &si);
  if (rc != MC_OK) return rc;
  const int C = si.channels;
  const bool narrow = si.bits <= 16;
  int16_t* o16 = static_cast<int16_t*>(out);
  int32_t* o32 = static_cast<int32_t*>(out);
  std::vector<int32_t> buf;
  int64_t pos = si.first_frame, written = 0;
  while (pos + 2 <= n) {
    if (!(data[pos] == 0xFF && (data[pos + 1] & 0xFE) == 0xF8)) {
      if (si.total > 0 && written >= si.total) break;   // trailing bytes (e.g. an ID3v1 tag)
      return MC_ERR_ARG;
    }
    BitReader br(data, n, pos);
    br.bits(14); br.bit();
    br.bit();                                    // blocking strategy: the coded number is not needed for sequential decode
    const int bs_code = static_cast<int>(br.bits(4));
    const int sr_code = static_cast<int>(br.bits(4));
    const int ch_code = static_cast<int>(br.bits(4));
    const int sz_code = static_cast<int>(br.bits(3));
    if (br.bit() != 0) return MC_ERR_ARG;
    {                                            // UTF-8 style coded frame / sample number
      const uint32_t first = static_cast<uint32_t>(br.bits(8));
      int extra = 0;
      if (first & 0x80) { uint32_t m = 0x40; while (first & m) { ++extra; m >>= 1; } if (extra == 0 || extra > 6) return MC_ERR_ARG; }
      for (int i = 0; i < extra; ++i) if ((br.bits(8) & 0xC0) != 0x80) return MC_ERR_ARG;
    }
    int blocksize;
    if (bs_code == 0) return MC_ERR_ARG;
    else if (bs_code == 1) blocksize = 192;
    else if (bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = static_cast<int>(br.bits(8)) + 1;
    else if (bs_code == 7) blocksize = static_cast<int>(br.bits(16)) + 1;
    else blocksize = 256 << (bs_code - 8);
    if (sr_code == 12) br.bits(8); else if (sr_code == 13 || sr_code == 14) br.bits(16); else if (sr_code == 15) return MC_ERR_ARG;
    static const int kBits[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    int bps = kBits[sz_code];
    if (bps < 0) return MC_ERR_ARG;
    if (bps == 0) bps = si.bits;
    if (bps != si.bits) return MC_ERR_ARG;       // mid-stream format changes are not supported
    const int64_t hdr_end = br.byte_pos();
    const uint8_t want8 = static_cast<uint8_t>(br.bits(8));
    if (br.fail || crc8(data + pos, hdr_end - pos) != want8) return MC_ERR_ARG;
    int nch;
    if (ch_code < 8) nch = ch_code + 1; else if (ch_code <= 10) nch = 2; else return MC_ERR_ARG;
    if (nch != C) return MC_ERR_ARG;
    buf.resize(static_cast<size_t>(nch) * blocksize);
    for (int c = 0; c < nch; ++c) {
      const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
      if (bps + (side ? 1 : 0) > 32) return MC_ERR_ARG;   // 33-bit side channels of 32-bit streams are not supported
      if (!read_subframe(br, buf.data() + static_cast<size_t>(c) * blocksize, blocksize, bps + (side ? 1 : 0))) return MC_ERR_ARG;
    }
    br.align();
    const int64_t body_end = br.byte_pos();
    const uint16_t want16 = static_cast<uint16_t>(br.bits(16));
    if (br.fail || crc16(data + pos, body_end - pos) != want16) return MC_ERR_ARG;
    int32_t* c0 = buf.data();
    int32_t* c1 = nch > 1 ? buf.data() + blocksize : nullptr;
    if (ch_code == 8) { for (int i = 0; i < blocksize; ++i) c1[i] = c0[i] - c1[i]; }
    else if (ch_code == 9) { for (int i = 0; i < blocksize; ++i) c0[i] = c0[i] + c1[i]; }
    else if (ch_code == 10) {
      for (int i = 0; i < blocksize; ++i) {
        const int32_t side = c1[i];
        const int32_t mid = static_cast<int32_t>((static_cast<uint32_t>(c0[i]) << 1) | (static_cast<uint32_t>(side) & 1u));
        c0[i] = (mid + side) >> 1;
        c1[i] = (mid - side) >> 1;
      }
    }
    int64_t take = blocksize;
    if (si.total > 0 && written + take > si.total) take = si.total - written;
    if (written + take > capacity_frames) return MC_ERR_NOMEM;
    const int up_shift = 32 - si.bits;
    for (int c = 0; c < nch; ++c) {
      const int32_t* src = buf.data() + static_cast<size_t>(c) * blocksize;
      if (narrow) { for (int64_t i = 0; i < take; ++i) o16[(written + i) * C + c] = static_cast<int16_t>(src[i] << (16 - si.bits)); }
      else { for (int64_t i = 0; i < take; ++i) o32[(written + i) * C + c] = static_cast<int32_t>(static_cast<uint32_t>(src[i]) << up_shift); }
    }
    written += take;
    pos = br.byte_pos();
  }
  *decoded = written;
  if (si.total > 0 && written != si.total) return MC_ERR_ARG;
  return MC_OK;
}

}  // extern "C"
