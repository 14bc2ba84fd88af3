// emu_fused.cpp — TEST INFRASTRUCTURE ONLY.
//
// The FUSED schedule of the product (k_fused_pvb + k_fused_adv per iteration, k_fused_dry for the
// dry sweep; csrc/wsb_fused_kernels.cuh, compiled here UNCHANGED with -DWSB_HOST_EMU) executed on
// the host by the little CUDA execution model of cuda_emu.h, with the launch sequence and the
// read-back views of csrc/wsb200.cu (fused_iteration, dry_iteration, wsb_read_rect) restated for a
// single domain.  tests/test_host_cells.py compares it with the oracle bit for bit: staging through
// (emulated) TMA boxes and through the register fallback on the outer ring of tiles, halos, the
// in-place stencil sweeps, the near back-trace and its hand-over to the exact path, the all-air
// shortcut, the folded pressure pass and the running maximum are then all checked on the CPU.
// Nothing in the product links or loads this file.
#include "cuda_emu.h"

#define WSB_HOST_EMU 1
#include "../../2d-weather-sandbox_b200/csrc/wsb_fused_kernels.cuh"
#include "../../2d-weather-sandbox_b200/csrc/wsb_particles.cuh"

using namespace wsb;

namespace {
struct Field {
  std::vector<float> data[4];
  Planes4 p;
  void alloc(size_t n) {
    for (int k = 0; k < 4; k++) { data[k].assign(n, 0.0f); p.c[k] = data[k].data(); }
  }
};
struct Sim {
  int W, H;  // W = columns of the local arrays (the pitch): the whole grid, or a strip with its ghost columns
  Geom g;
  DevParams dp;
  Field base[2], water[2], light[2];
  std::vector<int> wall[2];
  std::vector<float4> fb;
  std::vector<float2> dep;
  // sprite origins + dirty-tile map of the particle pass (csrc/wsb200.cu: alloc_all)
  std::vector<float4> org4;
  std::vector<float2> org2;
  std::vector<int> dirty, dirtyList;
  int dirtyCount[2] = {0, 0};
  SpriteGrid sg{};
  std::vector<unsigned char> tilewalls;  // k_fused_dry's tile map
  bool tilewalls_valid = false;
  std::vector<float> initial_T, sndT, sndW, sndV;
  unsigned maxv = 0;
  std::vector<float> drops[2];
  int ND = 0, last_drops = 0;
  float lightning[4] = {0.f, 0.f, 0.f, 0.f}, inactive = 0.0f;
  bool fb_dirty = false;
  bool even = true, pressure_pending = false, use_tma = false;
  long long iter = 0;
  long long launches = 0;
};
CUtensorMap map_of(const Sim& s, const void* plane, int bw, int bh) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  m.base = plane; m.W = s.W; m.H = s.H; m.boxW = bw; m.boxH = bh;
  return m;
}
GlobalCtx ctx(Sim& s, int b, int w, int wl, int l) {
  GlobalCtx c;
  c.base = s.base[b].p; c.water = s.water[w].p; c.wall = s.wall[wl].data(); c.vortf = nullptr; c.light = s.light[l].p;
  c.fb = s.fb.data(); c.dep = s.dep.data(); c.g = s.g;
  return c;
}
void derived(Sim& s) {
  s.dp.sinSun = (float)sin((double)s.dp.in.sunAngle);
  s.dp.cosSun = (float)cos((double)s.dp.in.sunAngle);
  s.dp.iterNum = (float)s.iter;
  s.dp.iterI = (int)s.dp.iterNum;
}
dim3 tile_grid(const Sim& s, int ty) { return dim3((s.W + kTX - 1) / kTX, (s.H + ty - 1) / ty); }

// csrc/wsb200.cu: fused_iteration, single domain, no particles
void fused_iteration(Sim& s) {
  s.tilewalls_valid = false;
  derived(s);
  const int src = s.even ? 0 : 1, dst = s.even ? 1 : 0;
  {
    GlobalCtx c = ctx(s, 1, 1, 1, 0);
    TileMaps<11> maps;
    for (int k = 0; k < 4; k++) {
      maps.m[k] = map_of(s, s.base[1].p.c[k], kSW1, kSH1);
      maps.m[5 + k] = map_of(s, s.water[1].p.c[k], kTX, kTY);
    }
    maps.m[4] = map_of(s, s.wall[1].data(), kSW1, kSH1);
    maps.m[9] = map_of(s, s.light[0].p.c[0], kTX, kTY);
    maps.m[10] = map_of(s, s.light[0].p.c[1], kTX, kTY);
    const DevParams d = s.dp;
    const int useTma = s.use_tma, applyPressure = s.pressure_pending ? 1 : 0, useFb = (s.fb_dirty && s.sg.dirty) ? 1 : 0;
    auto launch_pvb = [&](int cx0, int cx1, int gapAt, int gapLen) {
      c.g.cx0 = cx0; c.g.cx1 = cx1; c.g.cxGapAt = gapAt; c.g.cxGapLen = gapLen;
      emu::launch(dim3((cx1 - cx0 - gapLen + kTX - 1) / kTX, (s.H + kTY - 1) / kTY), kNT, kSmem1, [&] {
        k_fused_pvb(c, d, maps, useTma, s.initial_T.data(), applyPressure, useFb, s.fb.data(), s.dep.data(), s.sg, s.base[0].p, s.water[0].p,
                    s.wall[0].data());
      });
      s.launches++;
    };
    // a strip launches the tiles clear of the ghost columns first and the two edge tile columns after the ghost
    // exchange has landed (csrc/wsb200.cu); the exchange itself is done by the test between iterations
    const int kGhostCols = 8;
    const int innerEnd = kTX + ((s.W - (kGhostCols + kHX) - kTX) / kTX) * kTX;
    constexpr int kNoGap = 0x7fffffff;
    if (!s.g.wrap && s.pressure_pending && innerEnd > kTX) {
      launch_pvb(kTX, innerEnd, kNoGap, 0);
      launch_pvb(0, s.W, kTX, innerEnd - kTX);  // both edge tile columns in one launch
    } else {
      launch_pvb(0, s.W, kNoGap, 0);
    }
    s.fb_dirty = false;  // the kernel has consumed the feedback and zeroed the texels that were hit (app.js:5933-5934 folded in)
  }
  {
    GlobalCtx c = ctx(s, 0, 0, 0, src);
    TileMaps<12> maps;
    for (int k = 0; k < 4; k++) {
      maps.m[k] = map_of(s, s.base[0].p.c[k], kSW2, kSH2);
      maps.m[4 + k] = map_of(s, s.water[0].p.c[k], kSW2, kSH2);
    }
    maps.m[8] = map_of(s, s.wall[0].data(), kSW2, kSH2);
    maps.m[9] = map_of(s, s.light[src].p.c[0], kSW2, kSH2);
    maps.m[10] = map_of(s, s.light[src].p.c[2], kSW2, kSH2);
    maps.m[11] = map_of(s, s.light[src].p.c[3], kSW2, kSH2);
    const DevParams d = s.dp;
    const int useTma = s.use_tma;
    auto launch_adv = [&](int cx0, int cx1, int gapAt, int gapLen) {
      c.g.cx0 = cx0; c.g.cx1 = cx1; c.g.cxGapAt = gapAt; c.g.cxGapLen = gapLen;
      emu::launch(dim3((cx1 - cx0 - gapLen + kTX - 1) / kTX, (s.H + kTY - 1) / kTY), kNT, kSmem2, [&] {
        k_fused_adv(c, d, maps, useTma, s.initial_T.data(), s.sndT.data(), s.sndW.data(), s.sndV.data(), s.base[1].p, s.water[1].p,
                    s.wall[1].data(), s.light[dst].p, &s.maxv);
      });
      s.launches++;
    };
    // a strip with the peer transport: the edge tile columns (which produce the neighbours' ghost columns) first
    const int advEdgeStart = ((s.W - 2 * 8) / kTX) * kTX;
    if (!s.g.wrap && advEdgeStart > kTX) {
      launch_adv(0, s.W, kTX, advEdgeStart - kTX);
      launch_adv(kTX, advEdgeStart, 0x7fffffff, 0);
    } else {
      launch_adv(0, s.W, 0x7fffffff, 0);
    }
  }
  s.even = !s.even;
  s.pressure_pending = true;
  if (s.dp.p.enablePrecipitation && s.ND > 0) {  // csrc/wsb200.cu: precipitation()
    const int psrc = s.even ? 1 : 0, pdst = s.even ? 0 : 1;  // chosen before the lighting block toggled `even`
    derived(s);
    const DevParams d = s.dp;
    emu::launch(dim3((s.ND + 255) / 256, 1), 256, 0, [&] {
      k_precipitation(s.drops[psrc].data(), s.drops[pdst].data(), s.base[1].p, s.water[1].p, s.fb.data(), s.dep.data(), s.sg, s.lightning,
                      &s.inactive, s.g, d, s.ND);
    });
    const dim3 ctas(std::min(5, s.sg.tilesX * s.sg.tilesY), 1);  // persistent grid walking the dirty-tile list
    emu::launch(ctas, 256, kSmemBox, [&] { k_boxsum(s.sg, s.fb.data(), s.dep.data(), s.W, s.H, s.W); });
    emu::launch(ctas, 256, 0, [&] { k_clear_origins(s.sg, s.W, s.H, reinterpret_cast<unsigned*>(&s.dirtyCount[1])); });
    s.fb_dirty = true;
    s.last_drops = pdst;
    emu::launch(dim3(1, 1), 32, 0, [&] { k_latch(s.fb.data(), &s.inactive, s.lightning, d.iterNum, (s.iter % 600 == 0) ? 1 : 0); });
    s.launches += 4;
  }
  s.iter++;
}
// csrc/wsb200.cu: dry_iteration, FUSED schedule
void dry_iteration(Sim& s) {
  derived(s);
  GlobalCtx c = ctx(s, 1, 1, 1, 0);
  TileMaps<5> maps;
  for (int k = 0; k < 4; k++) maps.m[k] = map_of(s, s.base[1].p.c[k], kSWD, kSHD);
  maps.m[4] = map_of(s, s.wall[1].data(), kSWD, kSHD);
  const DevParams d = s.dp;
  const int useTma = s.use_tma, applyPressure = s.pressure_pending ? 1 : 0;
  const dim3 tiles = tile_grid(s, kTYD);
  if (!s.tilewalls_valid) {  // csrc/wsb200.cu: dry_iteration — the wall texture may have changed since the last dry sweep
    s.tilewalls.assign((size_t)tiles.x * tiles.y, 0xCD);
    emu::launch(tiles, 256, 0, [&] { k_wall_tilemap(c, s.tilewalls.data()); });
    s.tilewalls_valid = true;
    s.launches++;
  }
  emu::launch(tiles, kNT, kSmemDry, [&] { k_fused_dry(c, d, maps, useTma, applyPressure, s.tilewalls.data(), s.base[0].p, &s.maxv); });
  s.launches++;
  std::swap(s.base[0], s.base[1]);
  s.pressure_pending = true;
  s.iter++;
}
}  // namespace

extern "C" {
// one x-strip of a Wg-wide periodic grid: owned global columns [x_begin, x_begin + lw), `ghost` extra columns per
// side (wsb_create with n_ranks > 1); ghost = 0 and lw = Wg is the single domain
void* ef_create_strip(int Wg, int H, int x_begin, int lw, int ghost) {
  Sim* s = new Sim();
  const int W = lw + 2 * ghost;
  s->W = W; s->H = H;
  const size_t n = (size_t)W * H;
  for (int k = 0; k < 2; k++) { s->base[k].alloc(n); s->water[k].alloc(n); s->light[k].alloc(n); s->wall[k].assign(n, 0); }
  s->fb.assign(n, make_float4(0.f, 0.f, 0.f, 0.f)); s->dep.assign(n, make_float2(0.f, 0.f));
  s->initial_T.assign(H + 2, 0.0f); s->sndT.assign(H + 2, 0.0f); s->sndW.assign(H + 2, 0.0f); s->sndV.assign(H + 2, 0.0f);
  Geom& g = s->g;  // as wsb_create
  g.Wg = Wg; g.H = H; g.pitch = W; g.gx0 = x_begin - ghost; g.wrap = ghost == 0 ? 1 : 0; g.cx0 = 0; g.cx1 = W;
  g.cxGapAt = 0x7fffffff; g.cxGapLen = 0; g.ox0 = ghost; g.ox1 = ghost + lw;
  g.texelX = (float)(1.0 / (double)Wg); g.texelY = (float)(1.0 / (double)H);
  g.Wf = (float)Wg; g.Hf = (float)H;
  g.ltexelX = 1.0f / g.Wf; g.ltexelY = 1.0f / g.Hf;
  g.cellHeightComp = 300.0f / g.Hf;
  g.nearV = 0.9f;
  // alloc_all: TMA needs 16-byte row strides and at least one box per dimension
  s->use_tma = W % 4 == 0 && W >= kSW1 && H >= (kSH1 > kSHD ? kSH1 : kSHD);
  memset(&s->dp, 0, sizeof(s->dp));
  s->dp.in.userInputType = -1;
  derived(*s);
  return s;
}
void* ef_create(int W, int H) { return ef_create_strip(W, H, 0, W, 0); }
void ef_destroy(void* h) { delete (Sim*)h; }
void ef_upload(void* h, const float* base, const float* water, const int8_t* wall) {
  Sim& s = *(Sim*)h;
  const size_t n = (size_t)s.W * s.H;
  for (int k = 0; k < 2; k++)
    for (size_t i = 0; i < n; i++) {
      for (int ch = 0; ch < 4; ch++) { s.base[k].data[ch][i] = base[i * 4 + ch]; s.water[k].data[ch][i] = water[i * 4 + ch]; s.light[k].data[ch][i] = 0.0f; }
      int w; memcpy(&w, wall + i * 4, 4); s.wall[k][i] = w;
    }
  s.even = true; s.iter = 0; s.pressure_pending = false; s.maxv = 0; s.tilewalls_valid = false;
  derived(s);
}
void ef_upload_drops(void* h, const float* drops, int nd) {
  Sim& s = *(Sim*)h;
  s.ND = nd;
  for (int k = 0; k < 2; k++) s.drops[k].assign(drops, drops + (size_t)nd * 5);
  s.last_drops = 0; s.inactive = 0.0f; s.fb_dirty = false;
  memset(s.lightning, 0, sizeof(s.lightning));
  std::fill(s.fb.begin(), s.fb.end(), make_float4(0.f, 0.f, 0.f, 0.f));
  std::fill(s.dep.begin(), s.dep.end(), make_float2(0.f, 0.f));
  // csrc/wsb200.cu: alloc_all / zero_transients for n_droplets > 0
  SpriteGrid& sg = s.sg;
  sg.Po = s.W + 1;
  sg.tilesX = (s.W + kPTX - 1) / kPTX;
  sg.tilesY = (s.H + kPTY - 1) / kPTY;
  s.org4.assign((size_t)sg.Po * (s.H + 1), make_float4(0.f, 0.f, 0.f, 0.f));
  s.org2.assign((size_t)sg.Po * (s.H + 1), make_float2(0.f, 0.f));
  s.dirty.assign((size_t)sg.tilesX * sg.tilesY, 0);
  s.dirtyList.assign((size_t)sg.tilesX * sg.tilesY, -1);
  s.dirtyCount[0] = s.dirtyCount[1] = 0;
  sg.org4 = s.org4.data(); sg.org2 = s.org2.data(); sg.dirty = s.dirty.data(); sg.dirtyList = s.dirtyList.data(); sg.dirtyCount = s.dirtyCount;
}
// non-zero sprite-origin cells left after a step (k_clear_origins must leave none), and dirty tiles still flagged
int ef_origin_residue(void* h, int* dirty_tiles) {
  Sim& s = *(Sim*)h;
  int n = 0, d = 0;
  for (const float4& v : s.org4) n += (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f);
  for (const float2& v : s.org2) n += (v.x != 0.f || v.y != 0.f);
  for (size_t i = 0; i < s.dirty.size(); i++) d += s.dirty[i] != 0;
  if (s.dirtyCount[0] != 0 || s.dirtyCount[1] != 0) n += 1000000;  // the list must have been reset
  *dirty_tiles = d;
  return n;
}
void ef_read_drops(void* h, float* dst) { Sim& s = *(Sim*)h; memcpy(dst, s.drops[s.last_drops].data(), (size_t)s.ND * 5 * 4); }
void ef_read_feedback(void* h, float* fb, float* dep) {
  Sim& s = *(Sim*)h;
  memcpy(fb, s.fb.data(), s.fb.size() * sizeof(float4));
  memcpy(dep, s.dep.data(), s.dep.size() * sizeof(float2));
}
void ef_get_latches(void* h, float* lightning4, float* inactive) { Sim& s = *(Sim*)h; memcpy(lightning4, s.lightning, 16); *inactive = s.inactive; }
void ef_set_inactive(void* h, float v) { ((Sim*)h)->inactive = v; }
void ef_set_params(void* h, const wsb_params* p) { ((Sim*)h)->dp.p = *p; }
void ef_set_frame_inputs(void* h, const wsb_frame_inputs* in) { Sim& s = *(Sim*)h; s.dp.in = *in; derived(s); }
void ef_set_profiles(void* h, const float* t0, const float* st, const float* sw, const float* sv) {
  Sim& s = *(Sim*)h;
  const size_t n = (size_t)s.H + 1;
  if (t0) memcpy(s.initial_T.data(), t0, n * 4);
  if (st) memcpy(s.sndT.data(), st, n * 4);
  if (sw) memcpy(s.sndW.data(), sw, n * 4);
  if (sv) memcpy(s.sndV.data(), sv, n * 4);
}
void ef_set_iter(void* h, long long it) { Sim& s = *(Sim*)h; s.iter = it; derived(s); }
int ef_uses_tma(void* h) { return ((Sim*)h)->use_tma ? 1 : 0; }
// the 13 planes whose ghost columns csrc/wsb200.cu exchanges after an iteration: base_1, water_1, wall_1 and the
// light texture just written (as 4-byte words)
void ef_exchange_planes(void* h, void** out) {
  Sim& s = *(Sim*)h;
  const int dst = s.even ? 0 : 1;  // `even` has been toggled since the lighting half of k_fused_adv wrote it
  int n = 0;
  for (int k = 0; k < 4; k++) out[n++] = s.base[1].p.c[k];
  for (int k = 0; k < 4; k++) out[n++] = s.water[1].p.c[k];
  out[n++] = s.wall[1].data();
  for (int k = 0; k < 4; k++) out[n++] = s.light[dst].p.c[k];
}
void ef_step(void* h, int n) { for (int i = 0; i < n; i++) fused_iteration(*(Sim*)h); }
void ef_step_dry(void* h, int n) { for (int i = 0; i < n; i++) dry_iteration(*(Sim*)h); }
float ef_max_velocity(void* h) { return __uint_as_float(((Sim*)h)->maxv); }
// wsb_read_rect (FUSED schedule), whole grid.  field: 0 base, 1 water, 2 wall, 3 light; view: 0 frameBuff_0, 1 frameBuff_1, 2 latest
void ef_read(void* h, int field, int view, void* dst) {
  Sim& s = *(Sim*)h;
  const size_t n = (size_t)s.W * s.H;
  const int v1 = view == 1 ? 1 : 0;
  if (field == 2) { memcpy(dst, s.wall[1].data(), n * 4); return; }
  float* o = (float*)dst;
  if (field == 0) {
    if (v1 == 0 && s.pressure_pending) {  // k_pressure_rect: the folded pressure pass applied on the fly
      GlobalCtx c = ctx(s, 1, 1, 1, 0);
      for (int y = 0; y < s.H; y++) for (int x = 0; x < s.W; x++) {
        float4 b = c.base4(x, y);
        char4 wYm = c.wall4(x, y - 1);
        pressure_cell(b.x, b.y, b.z, b.w, c.bx(x - 1, y), c.by(x, y - 1), c.bt(x, y - 1), wYm.x, wYm.y);
        float* q = o + ((size_t)y * s.W + x) * 4;
        q[0] = b.x; q[1] = b.y; q[2] = b.z; q[3] = b.w;
      }
      return;
    }
    for (size_t i = 0; i < n; i++) for (int ch = 0; ch < 4; ch++) o[i * 4 + ch] = s.base[1].data[ch][i];
    return;
  }
  Field& f = field == 1 ? s.water[v1] : (view == 2 ? s.light[s.even ? 0 : 1] : s.light[v1]);
  for (size_t i = 0; i < n; i++) for (int ch = 0; ch < 4; ch++) o[i * 4 + ch] = f.data[ch][i];
}
}  // extern "C"
