// glsl_shim.h — just enough of GLSL ES 3.00 in C++17 to compile the reference's OWN shader sources
// (read from /root/reference/shaders at build time, never copied into this repo) for the host.
// TEST INFRASTRUCTURE ONLY: part of the oracle/_ref build (oracle/ref_shim/build_ref.py); nothing in
// the product, and nothing that runs on the GPU box, includes or loads it.
//
// What this header decides is exactly what GLSL ES / WebGL leave to the implementation; every such
// choice is the canonical form of DESIGN.md "Spec freeze" (and therefore the oracle's and the
// kernels'): one fp32 rounding per written operation (build with -ffp-contract=off), min / max /
// clamp / mix / fract / mod / smoothstep / length / dot as their GLSL definitions, pow() with the
// constant exponents the shaders use as closed forms, sin / cos evaluated in double and rounded,
// NEAREST and LINEAR texture fetches, integer `% 0`.  Everything else — every line of shader logic —
// is the reference's text, compiled as it stands.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

// bound of the per-row uniform tables (reference: 126 vec4 = 504 rows, which caps its grids at 503 rows)
#define WSB_REF_PROFILE_VEC4S 1026
#define WSB_REF_MAX_ROWS (4 * WSB_REF_PROFILE_VEC4S - 1)

namespace glsl {

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec4;

// ---- swizzles: a view of components I... of a vector with S components, assignable ------------------
template <class V, class T, int S, int... I>
struct Swz {
  T d[S];
  operator V() const { return V(d[I]...); }
  Swz& operator=(const V& r) {
    const V c(r);  // a copy first: the source may alias this view
    int k = 0;
    ((d[I] = c.v[k++]), ...);
    return *this;
  }
  Swz& operator=(const Swz& r) { return *this = V(r); }
  template <int S2, int... J>
  Swz& operator=(const Swz<V, T, S2, J...>& r) { return *this = V(r); }
  Swz& operator+=(const V& r) { return *this = V(*this) + r; }
  Swz& operator-=(const V& r) { return *this = V(*this) - r; }
  Swz& operator*=(const V& r) { return *this = V(*this) * r; }
  Swz& operator/=(const V& r) { return *this = V(*this) / r; }
  Swz& operator+=(T r) { return *this = V(*this) + r; }
  Swz& operator-=(T r) { return *this = V(*this) - r; }
  Swz& operator*=(T r) { return *this = V(*this) * r; }
  Swz& operator/=(T r) { return *this = V(*this) / r; }
  T operator[](int i) const { return V(*this)[i]; }
};

// component index outside the vector is undefined in GLSL ES; canonical: clamped (the y-1
// sounding index at y = 0)
inline int clampi(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

#define GLSL_VEC_COMMON(V, T, N)                                   \
  V(const V& o) { for (int i = 0; i < N; i++) v[i] = o.v[i]; }     \
  V& operator=(const V& o) { for (int i = 0; i < N; i++) v[i] = o.v[i]; return *this; } \
  T& operator[](int i) { return v[clampi(i, N)]; }                 \
  const T& operator[](int i) const { return v[clampi(i, N)]; }

struct vec2 {
  union {
    float v[2];
    struct { float x, y; };
    struct { float r, g; };
    struct { float s, t; };
    Swz<vec2, float, 2, 0, 1> xy;
  };
  vec2() : v{0.0f, 0.0f} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic<A>::value>> explicit vec2(A a) : v{(float)a, (float)a} {}
  template <class A, class B> vec2(A a, B b) : v{(float)a, (float)b} {}
  explicit vec2(const ivec2& o);
  GLSL_VEC_COMMON(vec2, float, 2)
};
struct vec3 {
  union {
    float v[3];
    struct { float x, y, z; };
    struct { float r, g, b; };
    Swz<vec2, float, 3, 0, 1> xy;
    Swz<vec3, float, 3, 0, 1, 2> xyz, rgb;
    Swz<vec3, float, 3, 0, 0, 0> xxx;
    Swz<vec3, float, 3, 1, 1, 1> yyy;
    Swz<vec3, float, 3, 2, 2, 2> zzz;
  };
  vec3() : v{0.0f, 0.0f, 0.0f} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic<A>::value>> explicit vec3(A a) : v{(float)a, (float)a, (float)a} {}
  template <class A, class B, class C> vec3(A a, B b, C c) : v{(float)a, (float)b, (float)c} {}
  template <class C> vec3(const vec2& a, C c) : v{a.x, a.y, (float)c} {}
  GLSL_VEC_COMMON(vec3, float, 3)
};
struct vec4 {
  union {
    float v[4];
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
    Swz<vec2, float, 4, 0, 1> xy, rg;
    Swz<vec2, float, 4, 2, 3> zw, ba;
    Swz<vec3, float, 4, 0, 1, 2> xyz, rgb;
    Swz<vec3, float, 4, 0, 1, 3> xyw;
    Swz<vec3, float, 4, 0, 0, 0> xxx;
    Swz<vec3, float, 4, 1, 1, 1> yyy;
    Swz<vec3, float, 4, 2, 2, 2> zzz;
    Swz<vec3, float, 4, 3, 3, 3> www;
    Swz<vec4, float, 4, 0, 1, 2, 3> xyzw, rgba;
  };
  vec4() : v{0.0f, 0.0f, 0.0f, 0.0f} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic<A>::value>> explicit vec4(A a) : v{(float)a, (float)a, (float)a, (float)a} {}
  template <class A, class B, class C, class D> vec4(A a, B b, C c, D d) : v{(float)a, (float)b, (float)c, (float)d} {}
  template <class C, class D> vec4(const vec2& a, C c, D d) : v{a.x, a.y, (float)c, (float)d} {}
  vec4(const vec2& a, const vec2& b) : v{a.x, a.y, b.x, b.y} {}
  template <class D> vec4(const vec3& a, D d) : v{a.x, a.y, a.z, (float)d} {}
  GLSL_VEC_COMMON(vec4, float, 4)
};
struct ivec2 {
  union {
    int v[2];
    struct { int x, y; };
    Swz<ivec2, int, 2, 0, 1> xy;
  };
  ivec2() : v{0, 0} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic<A>::value>> explicit ivec2(A a) : v{(int)a, (int)a} {}
  template <class A, class B> ivec2(A a, B b) : v{(int)a, (int)b} {}
  explicit ivec2(const vec2& o) : v{(int)o.x, (int)o.y} {}
  GLSL_VEC_COMMON(ivec2, int, 2)
};
struct ivec4 {
  union {
    int v[4];
    struct { int x, y, z, w; };
    struct { int r, g, b, a; };
    Swz<ivec2, int, 4, 0, 1> xy;
    Swz<ivec2, int, 4, 2, 3> zw, ba;
  };
  ivec4() : v{0, 0, 0, 0} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic<A>::value>> explicit ivec4(A a) : v{(int)a, (int)a, (int)a, (int)a} {}
  template <class A, class B, class C, class D> ivec4(A a, B b, C c, D d) : v{(int)a, (int)b, (int)c, (int)d} {}
  GLSL_VEC_COMMON(ivec4, int, 4)
};
inline vec2::vec2(const ivec2& o) : v{(float)o.x, (float)o.y} {}

// ---- component-wise arithmetic (non-template on purpose: swizzle views convert implicitly) -----------
#define GLSL_BINOP(V, N, OP)                                                                         \
  inline V operator OP(const V& a, const V& b) { V r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] OP b.v[i]; return r; } \
  inline V operator OP(const V& a, float b) { V r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] OP b; return r; }         \
  inline V operator OP(float a, const V& b) { V r; for (int i = 0; i < N; i++) r.v[i] = a OP b.v[i]; return r; }         \
  inline V& operator OP##=(V& a, const V& b) { for (int i = 0; i < N; i++) a.v[i] = a.v[i] OP b.v[i]; return a; }        \
  inline V& operator OP##=(V& a, float b) { for (int i = 0; i < N; i++) a.v[i] = a.v[i] OP b; return a; }
#define GLSL_ALLOPS(V, N) GLSL_BINOP(V, N, +) GLSL_BINOP(V, N, -) GLSL_BINOP(V, N, *) GLSL_BINOP(V, N, /) \
  inline V operator-(const V& a) { V r; for (int i = 0; i < N; i++) r.v[i] = -a.v[i]; return r; }
GLSL_ALLOPS(vec2, 2)
GLSL_ALLOPS(vec3, 3)
GLSL_ALLOPS(vec4, 4)

// ---- built-ins, canonical forms ----------------------------------------------------------
inline float min(float a, float b) { return b < a ? b : a; }  // GLSL: (y < x) ? y : x
inline float max(float a, float b) { return a < b ? b : a; }  // GLSL: (x < y) ? y : x
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float abs(float x) { return fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float floor(float x) { return floorf(x); }
inline float fract(float x) { return x - floorf(x); }
inline float mod(float x, float y) { return x - y * floorf(x / y); }
inline float sqrt(float x) { return sqrtf(x); }
#ifdef WSB_REF_LIBM
inline float sin(float x) { return sinf(x); }
inline float cos(float x) { return cosf(x); }
#else
inline float sin(float x) { return (float)::sin((double)x); }  // evaluated in double, rounded once
inline float cos(float x) { return (float)::cos((double)x); }
#endif
inline float step(float e, float x) { return x < e ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x) {
  float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
// pow(m, 1/3): bit-level initial guess and four Newton steps, + - * / only
inline float cbrt_canon(float m) {
  if (!(m > 0.0f)) return 0.0f;
  uint u;
  memcpy(&u, &m, 4);
  u = u / 3u + 709921077u;
  float y;
  memcpy(&y, &u, 4);
  for (int k = 0; k < 4; k++) y = y - (y - m / (y * y)) * (1.0f / 3.0f);
  return y;
}
// pow with the constant exponents of the simulation shaders as closed forms; anything else: libm
inline float pow(float x, float y) {
#ifdef WSB_REF_LIBM  // sensitivity experiment only (profiles/tools/freeze_sensitivity.py): what another implementation's pow would give
  return powf(x, y);
#endif
  if (y == 17.0f) { float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8; return x16 * x; }
  if (y == 4.0f) { float x2 = x * x; return x2 * x2; }
  if (y == 2.0f) return x * x;
  if (y == 0.5f) return sqrtf(x);
  if (y == 1.0f / 3.0f) return cbrt_canon(x);
  return powf(x, y);
}
inline uint floatBitsToUint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; memcpy(&f, &u, 4); return f; }

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec2& a) { return sqrtf(a.x * a.x + a.y * a.y); }
inline float length(const vec3& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }

#define GLSL_MAP1(V, N, F) inline V F(const V& a) { V r; for (int i = 0; i < N; i++) r.v[i] = F(a.v[i]); return r; }
#define GLSL_VECFUNCS(V, N)                                                                                         \
  GLSL_MAP1(V, N, abs) GLSL_MAP1(V, N, floor) GLSL_MAP1(V, N, fract) GLSL_MAP1(V, N, sin) GLSL_MAP1(V, N, cos) GLSL_MAP1(V, N, sqrt) \
  inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; i++) r.v[i] = min(a.v[i], b.v[i]); return r; } \
  inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; i++) r.v[i] = max(a.v[i], b.v[i]); return r; } \
  inline V min(const V& a, float b) { V r; for (int i = 0; i < N; i++) r.v[i] = min(a.v[i], b); return r; }         \
  inline V max(const V& a, float b) { V r; for (int i = 0; i < N; i++) r.v[i] = max(a.v[i], b); return r; }         \
  inline V clamp(const V& a, float lo, float hi) { V r; for (int i = 0; i < N; i++) r.v[i] = clamp(a.v[i], lo, hi); return r; } \
  inline V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < N; i++) r.v[i] = mix(a.v[i], b.v[i], t); return r; } \
  inline V mod(const V& a, float b) { V r; for (int i = 0; i < N; i++) r.v[i] = mod(a.v[i], b); return r; }
GLSL_VECFUNCS(vec2, 2)
GLSL_VECFUNCS(vec3, 3)
GLSL_VECFUNCS(vec4, 4)

// integer `a % b` with b == 0 is undefined in GLSL ES 3.00 (boundaryShader.frag:460, growth rates above 100).
// Canonical: the tick does not fire, i.e. the remainder is non-zero.  The translator wraps every
// right operand of % in glsl_nz().
struct NzInt { int v; };
inline NzInt glsl_nz(int b) { return NzInt{b}; }
inline int operator%(int a, NzInt b) { return b.v == 0 ? 1 : a % b.v; }

// ---- textures ------------------------------------------------------------------------------------------
enum { GL_NEAREST = 0, GL_LINEAR = 1, GL_REPEAT = 0, GL_CLAMP_TO_EDGE = 1 };
inline int wrapi(int i, int n) {
  if (i >= 0 && i < n) return i;  // the common case: no integer division
  i %= n;
  return i < 0 ? i + n : i;
}
inline int wrap_mode(int i, int n, int mode) { return mode == GL_REPEAT ? wrapi(i, n) : (i < 0 ? 0 : (i >= n ? n - 1 : i)); }

struct Texture {  // what a texture object holds: storage, size, channels, sampling state (app.js:5189-5317)
  const float* f = nullptr;
  const int8_t* i8 = nullptr;
  int W = 0, H = 0, C = 4;
  int filter = GL_NEAREST, wrapS = GL_REPEAT, wrapT = GL_REPEAT;
};
struct sampler2D { const Texture* t = nullptr; };
struct isampler2D { const Texture* t = nullptr; };

inline vec4 fetch_f(const Texture& t, int ix, int iy) {  // missing channels read (0, 0, 1)
  const float* p = t.f + ((size_t)iy * t.W + ix) * t.C;
  return vec4(p[0], t.C > 1 ? p[1] : 0.0f, t.C > 2 ? p[2] : 0.0f, t.C > 3 ? p[3] : 1.0f);
}
// NEAREST: texel floor(u * size), fp32.  LINEAR: full-fp32 bilinear about (u * size - 0.5) — except that a coordinate
// which IS a texel-centre coordinate, bit for bit as simShader.vert forms them ((i + 0.5) * texelSize, or a
// neighbour's +- texelSize), fetches that texel: lightingShader.frag:88,119 read IR_DOWN / IR_UP of the cells above
// and below through the LINEAR light sampler, and on grids whose 1 / size is not exact in fp32 the round trip
// (i + 0.5) * (1 / size) * size misses i + 0.5 by an ulp, which no hardware (8-bit weights) can see.
inline bool is_centre_coord(float u, int i, int size) {
  const float t = (float)(1.0 / (double)size);  // the uniform texelSize (app.js:5436-5437: JS double -> uniform2f)
  const float c0 = ((float)i + 0.5f) * t, cm = ((float)(i - 1) + 0.5f) * t, cp = ((float)(i + 1) + 0.5f) * t;
  return u == c0 || u == cm + t || u == cp + (-t);
}
inline vec4 texture(sampler2D s, const vec2& uv) {
  const Texture& t = *s.t;
  const float px = uv.x * (float)t.W, py = uv.y * (float)t.H;
  if (t.filter == GL_NEAREST)
    return fetch_f(t, wrap_mode((int)floorf(px), t.W, t.wrapS), wrap_mode((int)floorf(py), t.H, t.wrapT));
  float stx = px - 0.5f, sty = py - 0.5f;
  if (is_centre_coord(uv.x, (int)rintf(stx), t.W)) stx = rintf(stx);
  if (is_centre_coord(uv.y, (int)rintf(sty), t.H)) sty = rintf(sty);
  const float flx = floorf(stx), fly = floorf(sty);
  const float fx = stx - flx, fy = sty - fly;
  const int x0 = wrap_mode((int)flx, t.W, t.wrapS), x1 = wrap_mode((int)flx + 1, t.W, t.wrapS);
  const int y0 = wrap_mode((int)fly, t.H, t.wrapT), y1 = wrap_mode((int)fly + 1, t.H, t.wrapT);
  return mix(mix(fetch_f(t, x0, y0), fetch_f(t, x1, y0), fx), mix(fetch_f(t, x0, y1), fetch_f(t, x1, y1), fx), fy);
}
inline ivec4 texture(isampler2D s, const vec2& uv) {
  const Texture& t = *s.t;
  const int ix = wrap_mode((int)floorf(uv.x * (float)t.W), t.W, t.wrapS), iy = wrap_mode((int)floorf(uv.y * (float)t.H), t.H, t.wrapT);
  const int8_t* p = t.i8 + ((size_t)iy * t.W + ix) * 4;
  return ivec4(p[0], p[1], p[2], p[3]);
}
inline vec4 texelFetch(sampler2D s, const ivec2& p, int) { return fetch_f(*s.t, p.x, p.y); }

// `discard` inside main(): the fragment writes nothing
#define discard do { glsl_discarded = true; return; } while (0)

}  // namespace glsl
