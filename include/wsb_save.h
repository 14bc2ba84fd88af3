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

#ifdef __cplusplus
}
#endif
#endif
