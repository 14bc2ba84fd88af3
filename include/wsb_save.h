/*
 * wsb_save.h — native codec of the `.weathersandbox` container (reference: loadData app.js:1256-1366,
 * prepareDownload app.js:6575-6628; the reference uses pako, a JavaScript zlib, on one thread).
 *
 *   file   = u32 LE version id (263574036) | zlib stream
 *   stream = deflate of { u16 W, u16 H, base f32[H][W][4], water f32[H][W][4], wall i8[H][W][4],
 *                         droplets f32[N][5], i16 nStations, i16 xy[n][2], settings JSON }
 *
 * At 16384 x 4096 the payload is 2.4 GB: compression is split into independent chunks deflated on
 * several threads (each primed with the previous 32 KiB as dictionary, flushed to a byte boundary)
 * and stitched into ONE standard zlib stream, so pako.inflate / zlib.decompress read it unchanged.
 * CPU only (libwsbsave.so links zlib); no CUDA involved.
 */
#ifndef WSB_SAVE_H
#define WSB_SAVE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Upper bound of the compressed size of n payload bytes (for sizing the output buffer). */
int64_t wsb_save_compress_bound(int64_t n);

/* Deflate `payload` into a zlib stream at `out`.  level 0..9 (pako's default is 6), n_threads <= 0
 * = all cores.  Returns the stream length, or a negative value on error (-1 bad argument, -2
 * output buffer too small, -3 zlib failure). */
int64_t wsb_save_compress(const uint8_t* payload, int64_t n, uint8_t* out, int64_t out_cap, int32_t level, int32_t n_threads);

/* Inflate a zlib stream.  Returns the payload length, or negative (-2: out_cap too small, in which
 * case nothing useful is in `out`; -3: corrupt stream). */
int64_t wsb_save_decompress(const uint8_t* z, int64_t zn, uint8_t* out, int64_t out_cap);

/* Payload length of a zlib stream (inflates into a scratch window; no output buffer needed). */
int64_t wsb_save_inflated_size(const uint8_t* z, int64_t zn);

/* --- the container itself (loadData app.js:1256-1366, prepareDownload app.js:6575-6628) ---------------------- */

#define WSB_SAVE_VERSION_ID 263574036u         /* app.js:345 */
#define WSB_SAVE_LEGACY_VERSION_ID 1939327491u /* app.js:1265: no stations, no settings */

/* Where the pieces of an INFLATED payload are (byte offsets into it). */
typedef struct wsb_save_layout {
  int32_t width, height;
  int64_t n_droplets;    /* W*H/25 (NUM_DROPLETS_DEVIDER, app.js:452,1282) */
  int64_t off_base;      /* f32[H][W][4] */
  int64_t off_water;     /* f32[H][W][4] */
  int64_t off_wall;      /* i8 [H][W][4] */
  int64_t off_droplets;  /* f32[N][5] */
  int64_t n_stations;    /* 0 for the legacy version */
  int64_t off_stations;  /* i16[n][2] */
  int64_t off_settings;  /* UTF-8 JSON of guiControls up to the end of the payload; -1 for the legacy version */
  int64_t settings_len;
} wsb_save_layout;

/* loadData: validate and locate.  Returns 0, or -1 bad argument, -4 unknown version id, -5 truncated payload
 * ('Incompatible file!', app.js:1349). */
int32_t wsb_save_parse(const uint8_t* payload, int64_t n, uint32_t version, wsb_save_layout* out);

/* prepareDownload: payload size, and assembly into `out` (returns the size written, or negative: -1 bad
 * argument, -2 out_cap too small, -4 unknown version).  stations / settings are ignored for the legacy version. */
int64_t wsb_save_payload_size(int32_t width, int32_t height, int32_t n_stations, int64_t settings_len, uint32_t version);
int64_t wsb_save_serialise(uint8_t* out, int64_t out_cap, int32_t width, int32_t height, const float* base, const float* water,
                           const int8_t* wall, const float* droplets, const int16_t* stations, int32_t n_stations,
                           const char* settings, int64_t settings_len, uint32_t version);

#ifdef __cplusplus
}
#endif
#endif
