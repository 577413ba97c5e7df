// ORACLE PINNING — test infrastructure only. The one hand-written layer under the transpiled reference code (oracle/ref_transpile/cs2cpp.py):
// lane semantics of System.Numerics.Vector<T> / Vector / MathF from the .NET 8 BCL (not under /root/reference; SURVEY.md §8c lists what the path
// relies on: lane-wise IEEE-754 fp32 + - * / sqrt min max abs floor, compares producing all-ones masks, bitwise select; no FMA contraction — this
// file and everything including it is compiled with -ffp-contract=off). One lane per Vector (Vector<float>.Count == 1): constraint lanes are
// independent in every function on the path, so a 1-wide evaluation is the per-lane result of any hardware width.
#pragma once
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace bepu_ref {

#define REF_UNTRANSPILED(what) (std::fprintf(stderr, "bepu_ref: %s was not transpiled\n", what), std::abort())

template <class T> struct Vector;
template <> struct Vector<int> {
    int32_t v;
    Vector() = default;
    Vector(int32_t x) : v(x) {}
    static constexpr int Count = 1;
    static Vector Zero() { return Vector(0); }
    static Vector One() { return Vector(1); }
    static Vector AllBitsSet() { return Vector(-1); }
    int32_t operator[](int) const { return v; }
};
template <> struct Vector<float> {
    float v;
    Vector() = default;
    Vector(float x) : v(x) {}
    static constexpr int Count = 1;
    static Vector Zero() { return Vector(0.0f); }
    static Vector One() { return Vector(1.0f); }
    float operator[](int) const { return v; }
};
using VF = Vector<float>;
using VI = Vector<int>;

inline VF operator+(VF a, VF b) { return VF(a.v + b.v); }
inline VF operator-(VF a, VF b) { return VF(a.v - b.v); }
inline VF operator*(VF a, VF b) { return VF(a.v * b.v); }
inline VF operator/(VF a, VF b) { return VF(a.v / b.v); }
inline VF operator*(VF a, float b) { return VF(a.v * b); }
inline VF operator*(float a, VF b) { return VF(a * b.v); }
inline VF operator/(VF a, float b) { return VF(a.v / b); }
inline VF operator-(VF a) { return VF(-a.v); }
inline VI operator&(VI a, VI b) { return VI(a.v & b.v); }
inline VI operator|(VI a, VI b) { return VI(a.v | b.v); }
inline VI operator^(VI a, VI b) { return VI(a.v ^ b.v); }
inline VI operator~(VI a) { return VI(~a.v); }
inline VI operator+(VI a, VI b) { return VI(a.v + b.v); }
inline VI operator-(VI a, VI b) { return VI(a.v - b.v); }
inline VI operator*(VI a, VI b) { return VI(a.v * b.v); }
inline VI operator-(VI a) { return VI(-a.v); }
// compound assignment for every type that has the binary operator (C# synthesises `a op= b` from `a = a op b`)
template <class A, class B> inline auto operator+=(A& a, const B& b) -> decltype(a = a + b) { return a = a + b; }
template <class A, class B> inline auto operator-=(A& a, const B& b) -> decltype(a = a - b) { return a = a - b; }
template <class A, class B> inline auto operator*=(A& a, const B& b) -> decltype(a = a * b) { return a = a * b; }
template <class A, class B> inline auto operator/=(A& a, const B& b) -> decltype(a = a / b) { return a = a / b; }
template <class A, class B> inline auto operator&=(A& a, const B& b) -> decltype(a = a & b) { return a = a & b; }
template <class A, class B> inline auto operator|=(A& a, const B& b) -> decltype(a = a | b) { return a = a | b; }

inline int32_t f2i(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline float i2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

// System.Numerics.Vector (static helpers)
struct VectorOps {
    static VF Abs(VF a) { return VF(std::fabs(a.v)); }
    static VF SquareRoot(VF a) { return VF(std::sqrt(a.v)); }
    // Vector.Min/Max on x86 lower to minps/maxps: (a < b) ? a : b and (a > b) ? a : b. NaNs do not occur on the path.
    static VF Min(VF a, VF b) { return VF(a.v < b.v ? a.v : b.v); }
    static VF Max(VF a, VF b) { return VF(a.v > b.v ? a.v : b.v); }
    static VI Min(VI a, VI b) { return VI(a.v < b.v ? a.v : b.v); }
    static VI Max(VI a, VI b) { return VI(a.v > b.v ? a.v : b.v); }
    static VF Negate(VF a) { return VF(-a.v); }
    static VF Floor(VF a) { return VF(std::floor(a.v)); }
    static VI LessThan(VF a, VF b) { return VI(a.v < b.v ? -1 : 0); }
    static VI LessThanOrEqual(VF a, VF b) { return VI(a.v <= b.v ? -1 : 0); }
    static VI GreaterThan(VF a, VF b) { return VI(a.v > b.v ? -1 : 0); }
    static VI GreaterThanOrEqual(VF a, VF b) { return VI(a.v >= b.v ? -1 : 0); }
    static VI Equals(VF a, VF b) { return VI(a.v == b.v ? -1 : 0); }
    static VI Equals(VI a, VI b) { return VI(a.v == b.v ? -1 : 0); }
    static VI LessThan(VI a, VI b) { return VI(a.v < b.v ? -1 : 0); }
    static VI GreaterThan(VI a, VI b) { return VI(a.v > b.v ? -1 : 0); }
    static VF ConditionalSelect(VI c, VF a, VF b) { return VF(i2f((f2i(a.v) & c.v) | (f2i(b.v) & ~c.v))); }
    static VI ConditionalSelect(VI c, VI a, VI b) { return VI((a.v & c.v) | (b.v & ~c.v)); }
    static VI BitwiseAnd(VI a, VI b) { return VI(a.v & b.v); }
    static VI BitwiseOr(VI a, VI b) { return VI(a.v | b.v); }
    static VI AndNot(VI a, VI b) { return VI(a.v & ~b.v); }
    static VI OnesComplement(VI a) { return VI(~a.v); }
    static VF BitwiseAnd(VF a, VF b) { return VF(i2f(f2i(a.v) & f2i(b.v))); }
    static VF AndNot(VF a, VF b) { return VF(i2f(f2i(a.v) & ~f2i(b.v))); }
    static VF AsVectorSingle(VI a) { return VF(i2f(a.v)); }
    static VI AsVectorInt32(VF a) { return VI(f2i(a.v)); }
    static VF ConvertToSingle(VI a) { return VF((float)a.v); }
    static VI ConvertToInt32(VF a) { return VI((int32_t)a.v); }
    static bool LessThanAny(VF a, VF b) { return a.v < b.v; }
    static bool LessThanAny(VI a, VI b) { return a.v < b.v; }
    static bool LessThanAll(VF a, VF b) { return a.v < b.v; }
    static bool GreaterThanAny(VF a, VF b) { return a.v > b.v; }
    static bool EqualsAny(VI a, VI b) { return a.v == b.v; }
    static bool EqualsAll(VI a, VI b) { return a.v == b.v; }
};

// System.MathF / System.Math (scalar)
struct MathF {
    static constexpr float PI = 3.14159265358979323846f;
    static float Sqrt(float x) { return std::sqrt(x); }
    static float Abs(float x) { return std::fabs(x); }
    static float Min(float a, float b) { return a < b ? a : b; }
    static float Max(float a, float b) { return a > b ? a : b; }
    static float Sin(float x) { return std::sin(x); }
    static float Cos(float x) { return std::cos(x); }
    static float Pow(float a, float b) { return std::pow(a, b); }
    static float Floor(float x) { return std::floor(x); }
};
struct Math {
    static constexpr double PI = 3.14159265358979323846;
    static double Sqrt(double x) { return std::sqrt(x); }
    static double Abs(double x) { return std::fabs(x); }
    static float Abs(float x) { return std::fabs(x); }
    static float Min(float a, float b) { return a < b ? a : b; }
    static float Max(float a, float b) { return a > b ? a : b; }
    static int Min(int a, int b) { return a < b ? a : b; }
    static int Max(int a, int b) { return a > b ? a : b; }
};

}  // namespace bepu_ref
