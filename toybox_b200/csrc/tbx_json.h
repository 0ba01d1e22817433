/* tbx_json.h -- minimal JSON document model, parser and writer for the state/config codecs.
 *
 * Replaces the serde_json layer behind ctoybox's state_to_json / state_from_json / simulator_to_json
 * (reference call sites toybox/interventions/base.py:390-391,402-406).  Integers keep full 64-bit
 * range (rand.state holds u64 words); doubles are written with 17 significant digits so every f64
 * round-trips bit-exactly.
 */
#ifndef TBX_JSON_H
#define TBX_JSON_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <string>
#include <utility>
#include <vector>
#include <stdexcept>

namespace tbxjson {

struct Value;
typedef std::vector<std::pair<std::string, Value> > Members;

struct Value {
  enum Kind { Null, Bool, Int, UInt, Double, String, Array, Object } kind;
  bool b;
  int64_t i;   /* Int */
  uint64_t u;  /* UInt (only when > INT64_MAX) */
  double d;
  std::string s;
  std::vector<Value> a;
  Members o;
  Value() : kind(Null), b(false), i(0), u(0), d(0) {}
  static Value boolean(bool v) { Value x; x.kind = Bool; x.b = v; return x; }
  static Value integer(int64_t v) { Value x; x.kind = Int; x.i = v; return x; }
  static Value uinteger(uint64_t v) { Value x; if (v <= (uint64_t)INT64_MAX) { x.kind = Int; x.i = (int64_t)v; } else { x.kind = UInt; x.u = v; } return x; }
  static Value number(double v) { Value x; x.kind = Double; x.d = v; return x; }
  static Value string(const std::string &v) { Value x; x.kind = String; x.s = v; return x; }
  static Value array() { Value x; x.kind = Array; return x; }
  static Value object() { Value x; x.kind = Object; return x; }
  Value &set(const std::string &k, const Value &v) { o.push_back(std::make_pair(k, v)); return *this; }
  Value &push(const Value &v) { a.push_back(v); return *this; }

  bool is_null() const { return kind == Null; }
  const Value *find(const char *k) const {
    if (kind != Object) return 0;
    for (size_t n = 0; n < o.size(); n++) if (o[n].first == k) return &o[n].second;
    return 0;
  }
  bool has(const char *k) const { return find(k) != 0; }
  const Value &at(const char *k) const {
    const Value *v = find(k);
    if (!v) throw std::runtime_error(std::string("missing key '") + k + "'");
    return *v;
  }
  const Value &at(size_t n) const {
    if (kind != Array || n >= a.size()) throw std::runtime_error("array index out of range");
    return a[n];
  }
  size_t size() const { return kind == Array ? a.size() : kind == Object ? o.size() : 0; }
  void need(Kind k, const char *what) const { if (kind != k) throw std::runtime_error(std::string("expected ") + what); }
  bool as_bool() const {
    if (kind == Bool) return b;
    if (kind == Null) return false;
    if (kind == Int) return i != 0;
    throw std::runtime_error("expected a boolean");
  }
  int64_t as_i64() const {
    if (kind == Int) return i;
    if (kind == Bool) return b ? 1 : 0;
    if (kind == Double && d == std::floor(d) && std::fabs(d) < 9.0e18) return (int64_t)d;
    throw std::runtime_error("expected an integer");
  }
  int32_t as_i32() const {
    int64_t v = as_i64();
    if (v < INT32_MIN || v > INT32_MAX) throw std::runtime_error("integer out of i32 range");
    return (int32_t)v;
  }
  uint64_t as_u64() const {
    if (kind == UInt) return u;
    if (kind == Int && i >= 0) return (uint64_t)i;
    throw std::runtime_error("expected an unsigned integer");
  }
  double as_f64() const {
    if (kind == Double) return d;
    if (kind == Int) return (double)i;
    if (kind == UInt) return (double)u;
    throw std::runtime_error("expected a number");
  }
  const std::string &as_str() const { need(String, "a string"); return s; }
};

/* ---------------------------------------------------------------- parser */
struct Parser {
  const char *p, *end;
  Parser(const char *text, size_t n) : p(text), end(text + n) {}
  [[noreturn]] void fail(const char *msg) { throw std::runtime_error(std::string("JSON parse error: ") + msg); }
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool lit(const char *w) { size_t n = strlen(w); if ((size_t)(end - p) >= n && memcmp(p, w, n) == 0) { p += n; return true; } return false; }
  Value parse_value(int depth) {
    if (depth > 64) fail("nesting too deep");
    ws();
    if (p >= end) fail("unexpected end");
    char c = *p;
    if (c == '{') {
      p++;
      Value v = Value::object();
      ws();
      if (p < end && *p == '}') { p++; return v; }
      for (;;) {
        ws();
        if (p >= end || *p != '"') fail("expected a key");
        std::string k = parse_string();
        ws();
        if (p >= end || *p != ':') fail("expected ':'");
        p++;
        v.o.push_back(std::make_pair(k, parse_value(depth + 1)));
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; return v; }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      p++;
      Value v = Value::array();
      ws();
      if (p < end && *p == ']') { p++; return v; }
      for (;;) {
        v.a.push_back(parse_value(depth + 1));
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; return v; }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') return Value::string(parse_string());
    if (lit("true")) return Value::boolean(true);
    if (lit("false")) return Value::boolean(false);
    if (lit("null")) return Value();
    if (lit("NaN")) return Value::number(NAN);
    if (lit("Infinity")) return Value::number(INFINITY);
    if (lit("-Infinity")) return Value::number(-INFINITY);
    return parse_number();
  }
  std::string parse_string() {
    std::string out;
    p++; /* opening quote */
    while (p < end && *p != '"') {
      char c = *p++;
      if (c != '\\') { out.push_back(c); continue; }
      if (p >= end) fail("bad escape");
      char e = *p++;
      switch (e) {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
          if (end - p < 4) fail("bad \\u escape");
          unsigned cp = 0;
          for (int k = 0; k < 4; k++) {
            char h = *p++;
            cp = cp * 16 + (h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : 0);
          }
          if (cp < 0x80) out.push_back((char)cp);
          else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 63))); }
          else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 63))); out.push_back((char)(0x80 | (cp & 63))); }
          break;
        }
        default: fail("bad escape");
      }
    }
    if (p >= end) fail("unterminated string");
    p++;
    return out;
  }
  Value parse_number() {
    const char *s = p;
    bool neg = false, is_int = true;
    if (p < end && *p == '-') { neg = true; p++; }
    if (p >= end || !(*p >= '0' && *p <= '9')) fail("bad number");
    while (p < end && *p >= '0' && *p <= '9') p++;
    if (p < end && *p == '.') { is_int = false; p++; while (p < end && *p >= '0' && *p <= '9') p++; }
    if (p < end && (*p == 'e' || *p == 'E')) {
      is_int = false; p++;
      if (p < end && (*p == '+' || *p == '-')) p++;
      while (p < end && *p >= '0' && *p <= '9') p++;
    }
    std::string tok(s, p - s);
    if (is_int) {
      errno = 0;
      if (neg) {
        long long v = strtoll(tok.c_str(), 0, 10);
        if (errno == 0) return Value::integer(v);
      } else {
        unsigned long long v = strtoull(tok.c_str(), 0, 10);
        if (errno == 0) return Value::uinteger(v);
      }
    }
    return Value::number(strtod(tok.c_str(), 0));
  }
};

inline Value parse(const char *text) {
  if (!text) throw std::runtime_error("JSON parse error: null text");
  Parser ps(text, strlen(text));
  Value v = ps.parse_value(0);
  ps.ws();
  if (ps.p != ps.end) ps.fail("trailing characters");
  return v;
}

/* ---------------------------------------------------------------- writer */
inline void write_string(std::string &out, const std::string &s) {
  out.push_back('"');
  for (size_t n = 0; n < s.size(); n++) {
    unsigned char c = (unsigned char)s[n];
    if (c == '"') out += "\\\"";
    else if (c == '\\') out += "\\\\";
    else if (c == '\n') out += "\\n";
    else if (c == '\t') out += "\\t";
    else if (c == '\r') out += "\\r";
    else if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
    else out.push_back((char)c);
  }
  out.push_back('"');
}
inline void write(std::string &out, const Value &v) {
  char buf[40];
  switch (v.kind) {
    case Value::Null: out += "null"; break;
    case Value::Bool: out += v.b ? "true" : "false"; break;
    case Value::Int: snprintf(buf, sizeof buf, "%lld", (long long)v.i); out += buf; break;
    case Value::UInt: snprintf(buf, sizeof buf, "%llu", (unsigned long long)v.u); out += buf; break;
    case Value::Double:
      if (std::isnan(v.d)) out += "NaN";
      else if (std::isinf(v.d)) out += v.d > 0 ? "Infinity" : "-Infinity";
      else {
        snprintf(buf, sizeof buf, "%.17g", v.d);
        out += buf;
        if (!strpbrk(buf, ".eEn")) out += ".0"; /* keep it a float for the reader (serde writes 120.0) */
      }
      break;
    case Value::String: write_string(out, v.s); break;
    case Value::Array:
      out.push_back('[');
      for (size_t n = 0; n < v.a.size(); n++) { if (n) out.push_back(','); write(out, v.a[n]); }
      out.push_back(']');
      break;
    case Value::Object:
      out.push_back('{');
      for (size_t n = 0; n < v.o.size(); n++) {
        if (n) out.push_back(',');
        write_string(out, v.o[n].first);
        out.push_back(':');
        write(out, v.o[n].second);
      }
      out.push_back('}');
      break;
  }
}
inline std::string dump(const Value &v) { std::string s; write(s, v); return s; }

} /* namespace tbxjson */
#endif
