/*
 * wsb200.h — C ABI of libwsb200.so, the B200-native simulation core that replaces the
 * per-iteration simulation loop of 2D-Weather-Sandbox (reference: app.js:5830-6005 and the GLSL
 * passes it draws).  The reference has no plugin/FFI interface for this path: the boundary is the
 * set of WebGL interactions between app.js and the simulation state.  Each export below names the
 * reference interaction (file:line, relative to the reference checkout) it replaces one-for-one.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures; every function returns 0 on success, non-zero
 *     on failure with a message available from wsb_last_error() (the reference path has no error
 *     channel: GL errors are silent, shader failures throw strings at app.js:6644,6706).
 *   - the caller owns every host buffer; no pointer is retained after a call returns.
 *   - field layouts are exactly the reference's packed texture / save-file layouts:
 *       base  float[H][W][4]  (vx, vy, pressure, potential temperature)   RGBA32F  app.js:5122-5126
 *       water float[H][W][4]  (total, cloud, precip|soil moisture, smoke|snow)     app.js:5128-5133
 *       wall  int8 [H][W][4]  (type, distance, vertical distance, vegetation) RGBA8I app.js:5135-5139
 *       light float[H][W][4]  (sunlight, net heating, IR down, IR up)              app.js:5141-5145
 *       droplets float[N][5]  (x, y in [-1,1], water mass, ice mass, density)      app.js:4901-4913
 *     row 0 is the bottom row (GL origin), x is the fastest index.
 *   - all calls for one sim arrive from one host thread, in order (JS is single threaded).
 *     wsb_step is asynchronous (it only enqueues work, like the reference's draw calls); the read
 *     functions synchronise, like gl.readPixels.
 */
#ifndef WSB200_H
#define WSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSB_ABI_VERSION 1

typedef struct wsb_sim wsb_sim;

/* Size of an opaque NCCL bootstrap id (ncclUniqueId is 128 bytes). */
#define WSB_COMM_ID_BYTES 128

/* Replaces the allocation of the simulation textures / FBOs / droplet buffers
 * (app.js:5149-5317, 4885-5002).  For a multi-GPU run the grid is cut into x-strips, one process
 * per GPU: rank r owns global columns [x_begin, x_begin + local_width). */
typedef struct wsb_config {
  int32_t abi_version;   /* must be WSB_ABI_VERSION */
  int32_t width;         /* global sim_res_x */
  int32_t height;        /* sim_res_y */
  int32_t n_droplets;    /* NUM_DROPLETS (app.js:1282: W*H/25); 0 = no particle system */
  int32_t device;        /* CUDA device ordinal for this process */
  int32_t rank;          /* 0 .. n_ranks-1 */
  int32_t n_ranks;       /* 1 = single GPU */
  int32_t schedule;      /* WSB_SCHEDULE_* */
  uint8_t comm_id[WSB_COMM_ID_BYTES]; /* from wsb_comm_id_create on rank 0 (n_ranks > 1); all zero = no NCCL
                                         communicator (the strip must then be linked with wsb_connect_peers) */
} wsb_config;

enum {
  WSB_SCHEDULE_FUSED = 0,     /* product path: fused sm_100a kernels */
  WSB_SCHEDULE_REFERENCE = 1  /* one kernel per reference pass, same order as app.js:5830-6005
                                 (used for per-pass parity checks and as a perf baseline) */
};

/* Scalar uniforms of the simulation programs: static (app.js:5478-5640) and GUI-driven
 * (setGuiUniforms, app.js:3401-3443).  Names are the GLSL uniform names. */
typedef struct wsb_params {
  float dragMultiplier;          /* velocityShader.frag:16 */
  float wind;                    /* velocityShader.frag:18 */
  float vorticity;               /* boundaryShader.frag:26 */
  float landEvaporation;         /* boundaryShader.frag:28 */
  float waterEvaporation;        /* boundaryShader.frag:27 */
  float dynamicWaterTemperature; /* boundaryShader.frag:38 (0.0 or 1.0) */
  float evapHeat;                /* boundary/advection/precipitation */
  float waterWeight;             /* boundaryShader.frag:29 */
  float meltingHeat;             /* advectionShader.frag:35, precipitationShader.vert:38 */
  float condensationRate;        /* advectionShader.frag:36 */
  float globalDrying;            /* advectionShader.frag:40 */
  float globalHeating;           /* advectionShader.frag:41 */
  float soundingForcing;         /* advectionShader.frag:42 */
  float globalEffectsStartAlt;   /* already divided by simHeight (app.js:3425) */
  float globalEffectsEndAlt;     /* already divided by simHeight (app.js:3426) */
  float waterTemperature;        /* Kelvin (app.js:3427) */
  float greenhouseGases;         /* lightingShader.frag:27 */
  float waterGreenHouseEffect;   /* lightingShader.frag:28 */
  float IR_rate;                 /* lightingShader.frag:25 */
  float dryLapse;                /* app.js:5439 */
  float aboveZeroThreshold;      /* precipitationShader.vert:41-50 */
  float subZeroThreshold;
  float spawnChanceMult;         /* guiControls.spawnChance (app.js:3433) */
  float snowDensity;
  float fallSpeed;
  float growthRate0C;
  float growthRate_30C;
  float freezingRate;
  float meltingRate;
  float evapRate;
  int32_t enablePrecipitation;   /* guiControls.enablePrecipitation (app.js:5936) */
  int32_t reserved;
} wsb_params;

/* Per-frame uniforms: sun (updateSunlight, app.js:6557-6561), brush (app.js:5750-5808) and
 * airplane (app.js:3335).  They stay constant for all iterations of one wsb_step call, exactly
 * like the reference which sets them once per draw(). */
typedef struct wsb_frame_inputs {
  float sunAngle;           /* solar zenith angle in radians (app.js:6538-6539) */
  float sunIntensity;       /* W/m2 (app.js:6550) */
  float userInputValues[4]; /* x, y (normalised), intensity, brush size (advectionShader.frag:21) */
  float userInputMove[2];   /* advectionShader.frag:26 */
  int32_t userInputType;    /* advectionShader.frag:27; -1 = none */
  int32_t wrapHorizontally; /* advectionShader.frag:31 */
  float airplaneValues[4];  /* advectionShader.frag:29 */
} wsb_frame_inputs;

/* Fields readable through wsb_read_rect — the textures app.js reads with gl.readPixels. */
enum {
  WSB_FIELD_BASE = 0,       /* float4 */
  WSB_FIELD_WATER = 1,      /* float4 */
  WSB_FIELD_WALL = 2,       /* int8 x4 */
  WSB_FIELD_LIGHT = 3,      /* float4 */
  WSB_FIELD_FEEDBACK = 4,   /* float4 precipitationFeedbackTexture */
  WSB_FIELD_DEPOSITION = 5, /* float2 precipitationDepositionTexture */
  WSB_FIELD_CURL = 6,       /* float  (REFERENCE schedule only; the fused path never stores it) */
  WSB_FIELD_VORTFORCE = 7   /* float2 (REFERENCE schedule only) */
};

/* Which copy of a ping-pong pair a read returns. */
enum {
  WSB_VIEW_FRAMEBUFF_0 = 0, /* what the reference reads and saves: frameBuff_0 = base after the
                               pressure pass, water after the boundary pass, wall_0; light =
                               lightTexture_0 (app.js:6586-6593, 1082-1176) */
  WSB_VIEW_FRAMEBUFF_1 = 1, /* frameBuff_1 = base/water/wall after the advection pass; light =
                               lightTexture_1 */
  WSB_VIEW_LATEST = 2       /* light: the texture written by the last lighting pass; others as _0 */
};

/* --- lifecycle ------------------------------------------------------------------------------ */

/* Rank 0 of a multi-GPU job creates the bootstrap id and hands it to the other processes by any
 * host-side channel (the Python host uses torch.distributed broadcast). */
int wsb_comm_id_create(uint8_t out[WSB_COMM_ID_BYTES]);

/* Ghost-exchange transport of a multi-GPU run over peer memory (NVLink): every rank publishes a
 * window onto the planes its neighbours write ghost columns into (wsb_peer_info), the host
 * plumbing carries the blobs to the ring neighbours (torch.distributed all_gather in the Python
 * host), and wsb_connect_peers maps them (cudaIpcOpenMemHandle between processes, plain pointers
 * inside one process).  From then on every iteration's advection kernel is followed by ONE kernel
 * that stores this rank's outermost owned columns straight into the neighbours' ghost columns
 * and raises a sequence flag there; the next boundary kernel's edge tiles wait for the flag.
 * Without wsb_connect_peers a strip uses ncclSend/ncclRecv (needs a comm_id).  The reference is
 * single-GPU: no reference interaction is replaced. */
#define WSB_PEER_INFO_BYTES 256
int wsb_peer_info(wsb_sim* sim, uint8_t out[WSB_PEER_INFO_BYTES]);
int wsb_connect_peers(wsb_sim* sim, const uint8_t* left_info, const uint8_t* right_info);

/* A strip created with a comm_id AND linked with wsb_connect_peers has both transports; this selects the one the
 * following iterations use (every rank must make the same call between the same two iterations; results are
 * bit-identical either way).  Which one is faster depends on the strip width — measured on one NVSwitch box: peer at
 * 2 and 8 GPUs, NCCL at 4 (profiles/r3_multi_gpu.md) — so the Python host times both on the live state
 * (multi.calibrate_exchange).  Synchronises. */
/* WSB_EXCHANGE_PEER_COMPACT: the peer transport with a compact landing zone — the neighbours store the 8 ghost columns
 * of every plane and row into one contiguous 1.7-MB block per side of this rank's arena and the rank copies them into its
 * ghost columns itself; for rings with two distinct neighbours and wide strips, where stores scattered over every page of
 * two peer arenas are slow (profiles/r3_multi_gpu.md). */
enum { WSB_EXCHANGE_NCCL = 0, WSB_EXCHANGE_PEER = 1, WSB_EXCHANGE_PEER_COMPACT = 2 };
int wsb_set_exchange(wsb_sim* sim, int32_t transport);

/* app.js:5149-5317 + 4885-5002: allocate state; light, feedback, deposition, lightning, curl and
 * vortForce start zero-filled (texImage2D(..., null)); iterNum = 0, even = true. */
int wsb_create(const wsb_config* cfg, wsb_sim** out);

/* page unload */
int wsb_destroy(wsb_sim* sim);

/* app.js:5189-5234 (both ping-pong copies receive the same arrays) and 4917-4962 (both droplet
 * buffers).  Arrays are GLOBAL-size ([H][W][4] ...); every rank passes the same arrays and keeps
 * its own strip.  Also resets iterNum, even, light/feedback/lightning like a page load.
 * drops may be NULL when n_droplets == 0. Also used for the 'L' reload key (app.js:4628-4640). */
int wsb_upload(wsb_sim* sim, const float* base, const float* water, const int8_t* wall,
               const float* drops);

/* Multi-GPU variant of wsb_upload: arrays hold only this rank's PADDED strip,
 * [H][ghost + local_width + ghost][4], column i = global column (x_begin - ghost + i) mod W
 * (wsb_get_layout), so no rank has to materialise the whole grid. Single GPU: identical to
 * wsb_upload. */
int wsb_upload_local(wsb_sim* sim, const float* base, const float* water, const int8_t* wall,
                     const float* drops);

/* Strip geometry of this rank: owned columns [x_begin, x_begin + local_width) and the number of
 * ghost columns kept on each side (0 on a single GPU). */
int wsb_get_layout(wsb_sim* sim, int32_t* x_begin, int32_t* local_width, int32_t* ghost);

/* setGuiUniforms + the static uniforms (app.js:3401-3443, 5478-5640). */
int wsb_set_params(wsb_sim* sim, const wsb_params* p);

/* uniform arrays initial_Tv / realWorldSounding_{T,W,Vel}v (app.js:5444-5474, 5485-5502), here
 * generalised to length height+1 (the reference caps them at 504 entries).  Sounding arrays may
 * be NULL (= all zero, 'No valid sounding loaded', app.js:5462). */
int wsb_set_profiles(wsb_sim* sim, const float* initial_T, const float* sounding_T,
                     const float* sounding_W, const float* sounding_Vel);

/* updateSunlight / brush / airplane uniforms, once per frame. */
int wsb_set_frame_inputs(wsb_sim* sim, const wsb_frame_inputs* in);

/* --- the hot path --------------------------------------------------------------------------- */

/* The body of `for (i < IterPerFrame)` in draw() (app.js:5830-6005), n_iters times: velocity,
 * curl, vorticity, boundary, advection, pressure, lighting, feedback clear, precipitation
 * particles, inactive-droplet latch every 600 iterations (on device, no host sync), lightning
 * latch, iterNum++.  Asynchronous. */
int wsb_step(wsb_sim* sim, int32_t n_iters);

/* Block until everything enqueued so far has finished (gl.finish). */
int wsb_sync(wsb_sim* sim);

/* Run ONE reference pass (WSB_PASS_*) on the current buffers — REFERENCE schedule only; used by
 * the per-pass parity tests. */
enum {
  WSB_PASS_VELOCITY = 0, WSB_PASS_CURL = 1, WSB_PASS_VORTICITY = 2, WSB_PASS_BOUNDARY = 3,
  WSB_PASS_ADVECTION = 4, WSB_PASS_PRESSURE = 5, WSB_PASS_LIGHTING = 6, WSB_PASS_PRECIPITATION = 7,
  WSB_PASS_ITER_INC = 8,      /* iterNum++ (app.js:6002-6004) */
  WSB_PASS_ADVECTION_DRY = 9  /* advection of the base field only (the dry sweep's middle stage) */
};
int wsb_debug_run_pass(wsb_sim* sim, int32_t pass);

/* The "dry sweep" of BASELINE config 2: velocity -> advection of the base field -> pressure, fused
 * into one kernel, n_iters times.  Water, light and droplets are not touched. */
int wsb_step_dry(wsb_sim* sim, int32_t n_iters);

/* --- readbacks (gl.readPixels / getBufferSubData call sites) --------------------- */

/* Rectangle [x, x+w) x [y, y+h) in GLOBAL cell coordinates, no wrap (callers wrap themselves,
 * app.js:3061-3066).  A rank only fills the part of the rectangle inside its own strip and
 * leaves the rest of dst untouched; dst is a dense [h][w][channels] array. Synchronises. */
int wsb_read_rect(wsb_sim* sim, int32_t field, int32_t view, int32_t x, int32_t y, int32_t w,
                  int32_t h, void* dst);

/* Batched probes: n single texels of BASE / WATER / LIGHT at GLOBAL cell coordinates xy[n][2] in one
 * call and one synchronisation (the reference issues one gl.readPixels per weather station and
 * field every 208 iterations, app.js:1082-1176, 5997-6001 — each a pipeline stall).  dst is
 * float[n][4]; on a strip, points outside the rank's columns return zeros. */
int wsb_read_points(wsb_sim* sim, int32_t field, int32_t view, int32_t n, const int32_t* xy, float* dst);

/* app.js:5017-5019, 5084-5086, 6595-6597. buffer: 0/1 = precipVertexBuffer_0/_1, 2 = the one
 * written last. */
int wsb_read_droplets(wsb_sim* sim, int32_t buffer, int32_t first, int32_t count, float* dst);

/* app.js:5957-5967: the value latched into the `inactiveDroplets` uniform. */
int wsb_get_inactive_droplets(wsb_sim* sim, float* out);

/* app.js:5985-5994: the 1x1 lightningDataTexture. */
int wsb_get_lightning(wsb_sim* sim, float out[4]);

/* iterNum (app.js:440). */
int wsb_get_iter(wsb_sim* sim, int64_t* out);

/* Test / resume hook: set iterNum (the reference restarts it at 0 on every page load and never
 * saves it; the boundary pass keys its slow processes on iterNum % 100, % 600, app.js:5957). */
int wsb_set_iter(wsb_sim* sim, int64_t iter);

/* Local strip of this rank: [x_begin, x_begin + local_width). */
int wsb_get_strip(wsb_sim* sim, int32_t* x_begin, int32_t* local_width);

/* Largest |v| component the advection kernel has seen in this rank's OWN columns since upload
 * (cells / iteration).  The fused kernels stay exact for any value; multi-GPU strips are
 * bit-identical to the single-GPU run while it is <= 4 (ghost 8 = 3 + 1 + ceil|v|): every
 * synchronising call of a strip (wsb_sync, wsb_read_rect, wsb_last_step_ms) FAILS once it is
 * exceeded, instead of returning silently diverged fields. */
int wsb_get_max_velocity(wsb_sim* sim, float* out);

/* Debug aid: number of NaN / Inf values in the current base and water fields of this rank (the
 * reference has no failure detection beyond the local NaN-avoidance resets of advectionShader.frag:387-397).
 * Synchronises. */
int wsb_count_nonfinite(wsb_sim* sim, int64_t* out);

/* Number of kernels launched by this sim since creation (bench evidence). */
int wsb_get_launch_count(wsb_sim* sim, int64_t* out);

/* Milliseconds of device time between two internal CUDA events bracketing the most recent
 * wsb_step / wsb_step_dry call on the sim's stream (bench: the stream is library-owned, so
 * torch.cuda.Event cannot see it). Synchronises. */
int wsb_last_step_ms(wsb_sim* sim, float* out);

/* Per-kernel-class device time (bench / roofline evidence).  With profiling on, every launch of
 * wsb_step / wsb_step_dry is bracketed by CUDA events on the simulation's own stream;
 * wsb_kernel_time_ms returns the summed duration and the number of launches of one class during
 * the most recent step call. Synchronises. */
enum {
  WSB_KERNEL_PVB = 0,    /* k_fused_pvb: pressure -> velocity -> curl -> vorticity -> boundary */
  WSB_KERNEL_ADV = 1,    /* k_fused_adv: advection -> lighting */
  WSB_KERNEL_DRY = 2,    /* k_fused_dry: velocity -> advection(base) -> pressure */
  WSB_KERNEL_PRECIP = 3, /* the whole particle pass: k_precipitation + k_boxsum + k_clear_origins + k_latch */
  WSB_KERNEL_HALO = 4,   /* ghost exchange on the communication stream: k_push_ghosts (peer transport) or
                            pack + ncclSend/Recv + unpack (NCCL transport) */
  WSB_KERNEL_WAIT = 5,   /* k_wait_ghosts: time spent waiting for the neighbours' columns */
  WSB_KERNEL_EDGE = 6,   /* strips, peer transport: the edge tile columns of k_fused_pvb / k_fused_adv, which run on the
                            communication stream beside the interior tile columns (PVB / ADV then time the interior only) */
  WSB_KERNEL_SPRITES = 7 /* k_boxsum + k_clear_origins (part of PRECIP): the sprites as a box filter over the dirty tiles */
};
int wsb_set_profiling(wsb_sim* sim, int32_t on);
int wsb_kernel_time_ms(wsb_sim* sim, int32_t kernel, float* total_ms, int32_t* launches);

const char* wsb_last_error(void);
const char* wsb_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* WSB200_H */
