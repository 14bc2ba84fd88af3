// wsb_save.cpp — libwsbsave.so: multi-threaded zlib codec for .weathersandbox payloads (include/wsb_save.h).
#include "../../include/wsb_save.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {
constexpr int64_t kChunk = 4 << 20;  // payload bytes per independently deflated chunk
constexpr int kDict = 32768;         // deflate window: the previous chunk's tail primes the next one
}  // namespace

extern "C" {

int64_t wsb_save_compress_bound(int64_t n) {
  if (n < 0) return -1;
  const int64_t chunks = n / kChunk + 1;
  return 2 + (int64_t)deflateBound(nullptr, (uLong)std::min<int64_t>(n, kChunk)) * chunks + 16 * chunks + 4 + 64;
}

int64_t wsb_save_compress(const uint8_t* payload, int64_t n, uint8_t* out, int64_t out_cap, int32_t level, int32_t n_threads) {
  if (!payload || !out || n < 0 || level < 0 || level > 9) return -1;
  const int64_t nchunks = std::max<int64_t>(1, (n + kChunk - 1) / kChunk);
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, nchunks);
  std::vector<std::vector<uint8_t>> parts((size_t)nchunks);
  std::vector<uLong> adlers((size_t)nchunks);
  std::atomic<int64_t> next{0};
  std::atomic<int> failed{0};
  auto worker = [&]() {
    for (;;) {
      const int64_t c = next.fetch_add(1);
      if (c >= nchunks || failed.load()) return;
      const int64_t off = c * kChunk, len = std::min<int64_t>(kChunk, n - off);
      const bool last = c == nchunks - 1;
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { failed = 1; return; }
      if (c > 0) deflateSetDictionary(&zs, payload + off - kDict, kDict);  // kChunk >= kDict, so the tail exists
      std::vector<uint8_t>& buf = parts[(size_t)c];
      buf.resize(deflateBound(&zs, (uLong)len) + 16);
      zs.next_in = const_cast<Bytef*>(payload + off);
      zs.avail_in = (uInt)len;
      zs.next_out = buf.data();
      zs.avail_out = (uInt)buf.size();
      // non-final chunks end on a byte boundary with an empty stored block (Z_SYNC_FLUSH); the last one ends the stream
      const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
      if ((last && rc != Z_STREAM_END) || (!last && (rc != Z_OK || zs.avail_in != 0))) failed = 1;
      buf.resize(zs.total_out);
      deflateEnd(&zs);
      adlers[(size_t)c] = adler32(adler32(0L, Z_NULL, 0), payload + off, (uInt)len);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  if (failed.load()) return -3;
  int64_t total = 2 + 4;
  for (auto& p : parts) total += (int64_t)p.size();
  if (total > out_cap) return -2;
  uint8_t* o = out;
  *o++ = 0x78;  // CMF: deflate, 32 KiB window
  {             // FLG: level hint, no preset dictionary, (CMF*256 + FLG) % 31 == 0
    const int lvl = level < 2 ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3;
    int flg = lvl << 6;
    flg += 31 - ((0x78 * 256 + flg) % 31);
    *o++ = (uint8_t)flg;
  }
  uLong adler = adler32(0L, Z_NULL, 0);
  for (int64_t c = 0; c < nchunks; c++) {
    memcpy(o, parts[(size_t)c].data(), parts[(size_t)c].size());
    o += parts[(size_t)c].size();
    const int64_t len = std::min<int64_t>(kChunk, n - c * kChunk);
    adler = adler32_combine(adler, adlers[(size_t)c], (z_off_t)std::max<int64_t>(len, 0));
  }
  *o++ = (uint8_t)(adler >> 24);
  *o++ = (uint8_t)(adler >> 16);
  *o++ = (uint8_t)(adler >> 8);
  *o++ = (uint8_t)adler;
  return o - out;
}

static int64_t inflate_impl(const uint8_t* z, int64_t zn, uint8_t* out, int64_t out_cap, bool count_only) {
  if (!z || zn < 2) return -1;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit(&zs) != Z_OK) return -3;
  std::vector<uint8_t> scratch(count_only ? (1 << 20) : 0);
  Bytef overflow = 0;
  int64_t in_off = 0, total = 0;
  int rc = Z_OK;
  while (rc != Z_STREAM_END) {
    if (zs.avail_in == 0) {
      const int64_t take = std::min<int64_t>(zn - in_off, 1 << 30);
      if (take == 0) { inflateEnd(&zs); return -3; }
      zs.next_in = const_cast<Bytef*>(z + in_off);
      zs.avail_in = (uInt)take;
      in_off += take;
    }
    if (count_only) {
      zs.next_out = scratch.data();
      zs.avail_out = (uInt)scratch.size();
    } else {
      const int64_t room = std::min<int64_t>(out_cap - total, 1 << 30);
      if (room == 0) {  // the buffer is full: the stream may still hold its trailer, but no more data
        zs.next_out = &overflow;
        zs.avail_out = 1;
      } else {
        zs.next_out = out + total;
        zs.avail_out = (uInt)room;
      }
    }
    const uInt before = zs.avail_out;
    rc = inflate(&zs, Z_NO_FLUSH);
    if (!count_only && zs.next_out == &overflow + 1) { inflateEnd(&zs); return -2; }
    total += before - zs.avail_out;
    if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) { inflateEnd(&zs); return -3; }
  }
  inflateEnd(&zs);
  return total;
}

int64_t wsb_save_decompress(const uint8_t* z, int64_t zn, uint8_t* out, int64_t out_cap) {
  if (!out || out_cap < 0) return -1;
  return inflate_impl(z, zn, out, out_cap, false);
}

int64_t wsb_save_inflated_size(const uint8_t* z, int64_t zn) { return inflate_impl(z, zn, nullptr, 0, true); }

// --- container (loadData app.js:1256-1366, prepareDownload app.js:6575-6628) ---------------------------------
int32_t wsb_save_parse(const uint8_t* payload, int64_t n, uint32_t version, wsb_save_layout* out) {
  if (!payload || !out || n < 0) return -1;
  if (version != WSB_SAVE_VERSION_ID && version != WSB_SAVE_LEGACY_VERSION_ID) return -4;
  if (n < 4) return -5;
  wsb_save_layout L;
  memset(&L, 0, sizeof L);
  L.width = payload[0] | (payload[1] << 8);   // Uint16Array(buffer.slice(0, 4)), little endian
  L.height = payload[2] | (payload[3] << 8);
  const int64_t cells = (int64_t)L.width * L.height;
  L.n_droplets = cells / 25;
  L.off_base = 4;
  L.off_water = L.off_base + cells * 16;
  L.off_wall = L.off_water + cells * 16;
  L.off_droplets = L.off_wall + cells * 4;
  int64_t off = L.off_droplets + L.n_droplets * 20;
  if (n < off) return -5;
  L.off_settings = -1;
  if (version == WSB_SAVE_VERSION_ID) {
    if (n < off + 2) return -5;
    L.n_stations = (int16_t)(payload[off] | (payload[off + 1] << 8));
    if (L.n_stations < 0) return -5;
    L.off_stations = off + 2;
    off = L.off_stations + L.n_stations * 4;
    if (n < off) return -5;
    L.off_settings = off;
    L.settings_len = n - off;
  }
  *out = L;
  return 0;
}

int64_t wsb_save_payload_size(int32_t width, int32_t height, int32_t n_stations, int64_t settings_len, uint32_t version) {
  if (width <= 0 || height <= 0 || width > 65535 || height > 65535 || n_stations < 0 || settings_len < 0) return -1;
  if (version != WSB_SAVE_VERSION_ID && version != WSB_SAVE_LEGACY_VERSION_ID) return -4;
  const int64_t cells = (int64_t)width * height;
  int64_t n = 4 + cells * 36 + (cells / 25) * 20;
  if (version == WSB_SAVE_VERSION_ID) n += 2 + (int64_t)n_stations * 4 + settings_len;
  return n;
}

int64_t wsb_save_serialise(uint8_t* out, int64_t out_cap, int32_t width, int32_t height, const float* base, const float* water,
                           const int8_t* wall, const float* droplets, const int16_t* stations, int32_t n_stations,
                           const char* settings, int64_t settings_len, uint32_t version) {
  const int64_t need = wsb_save_payload_size(width, height, n_stations, settings_len, version);
  if (need < 0) return need;
  if (!out || !base || !water || !wall) return -1;
  if (out_cap < need) return -2;
  const int64_t cells = (int64_t)width * height, nd = cells / 25;
  if (nd > 0 && !droplets) return -1;
  uint8_t* p = out;
  p[0] = (uint8_t)(width & 0xff); p[1] = (uint8_t)(width >> 8); p[2] = (uint8_t)(height & 0xff); p[3] = (uint8_t)(height >> 8);
  p += 4;
  memcpy(p, base, (size_t)cells * 16); p += cells * 16;     // x86 / aarch64 hosts are little endian, like the typed arrays of the reference
  memcpy(p, water, (size_t)cells * 16); p += cells * 16;
  memcpy(p, wall, (size_t)cells * 4); p += cells * 4;
  if (nd) memcpy(p, droplets, (size_t)nd * 20);
  p += nd * 20;
  if (version == WSB_SAVE_VERSION_ID) {
    p[0] = (uint8_t)(n_stations & 0xff); p[1] = (uint8_t)((n_stations >> 8) & 0xff);   // Uint16Array.of(weatherStations.length)
    p += 2;
    if (n_stations) { if (!stations) return -1; memcpy(p, stations, (size_t)n_stations * 4); }
    p += (int64_t)n_stations * 4;
    if (settings_len) { if (!settings) return -1; memcpy(p, settings, (size_t)settings_len); }
    p += settings_len;
  }
  return p - out;
}

}  // extern "C"
