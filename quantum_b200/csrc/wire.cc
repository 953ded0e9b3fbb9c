// See wire.h. Field numbers / wire types: SURVEY.md Appendix A.
#include "wire.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

namespace tfqb {
namespace {

// ------------------------------------------------------------------ binary
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;

  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      const uint8_t b = *p++;
      if (shift < 64) v |= uint64_t(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  bool bytes(const uint8_t** s, size_t* n) {
    const uint64_t len = varint();
    if (!ok || len > size_t(end - p)) { ok = false; return false; }
    *s = p;
    *n = size_t(len);
    p += len;
    return true;
  }
  bool skip(int wt) {
    switch (wt) {
      case 0: varint(); return ok;
      case 1: if (end - p < 8) return ok = false; p += 8; return true;
      case 2: { const uint8_t* s; size_t n; return bytes(&s, &n); }
      case 5: if (end - p < 4) return ok = false; p += 4; return true;
      default: return ok = false;  // groups (3,4) and 6,7 are invalid here
    }
  }
  float fixed32f() {
    if (end - p < 4) { ok = false; return 0.f; }
    float f;
    std::memcpy(&f, p, 4);
    p += 4;
    return f;
  }
  double fixed64d() {
    if (end - p < 8) { ok = false; return 0.; }
    double d;
    std::memcpy(&d, p, 8);
    p += 8;
    return d;
  }
};

bool valid_utf8(const uint8_t* s, size_t n) {
  size_t i = 0;
  while (i < n) {
    const uint8_t c = s[i];
    int extra = c < 0x80 ? 0 : (c >> 5) == 6 ? 1 : (c >> 4) == 14 ? 2
                : (c >> 3) == 30 ? 3 : -1;
    if (extra < 0 || i + extra >= n + (extra == 0)) return false;
    for (int k = 1; k <= extra; ++k)
      if ((s[i + k] >> 6) != 2) return false;
    i += extra + 1;
  }
  return true;
}

// Iterate the fields of one message. `fn(field, wiretype, reader)` must
// consume the value and return false on error.
template <typename F>
bool each_field(const uint8_t* s, size_t n, F fn) {
  Reader r{s, s + n};
  while (!r.done()) {
    const uint64_t tag = r.varint();
    if (!r.ok) return false;
    const int field = int(tag >> 3), wt = int(tag & 7);
    if (field == 0) return false;
    if (!fn(field, wt, r) || !r.ok) return false;
  }
  return true;
}

bool read_string(Reader& r, int wt, std::string* out) {
  if (wt != 2) return false;
  const uint8_t* s; size_t n;
  if (!r.bytes(&s, &n) || !valid_utf8(s, n)) return false;
  out->assign(reinterpret_cast<const char*>(s), n);
  return true;
}

template <typename F>
bool read_msg(Reader& r, int wt, F fn) {
  if (wt != 2) return false;
  const uint8_t* s; size_t n;
  if (!r.bytes(&s, &n)) return false;
  return each_field(s, n, fn);
}

bool parse_arg_value(Reader& r, int wt, ArgPB* a) {
  // oneof: the last member on the wire wins and clears the others.
  return read_msg(r, wt, [&](int f, int w, Reader& rr) {
    switch (f) {
      case 1:
        if (w != 5) return false;
        a->float_value = rr.fixed32f();
        a->string_value.clear();
        a->has_bools = false;
        return true;
      case 3:
        a->float_value = 0.f;
        a->has_bools = false;
        return read_string(rr, w, &a->string_value);
      case 4:
        if (w != 1) return false;
        rr.fixed64d();          // double_value: never read by the C++ parser
        a->float_value = 0.f;
        a->string_value.clear();
        return true;
      case 2: {                 // bool_values (RepeatedBoolean)
        a->float_value = 0.f;
        a->string_value.clear();
        if (w != 2) return false;
        const uint8_t* s; size_t n;
        if (!rr.bytes(&s, &n)) return false;
        a->has_bools = true;
        a->bools_wire.assign(reinterpret_cast<const char*>(s), n);
        return true;
      }
      default: return rr.skip(w);
    }
  });
}

bool parse_arg(Reader& r, int wt, ArgPB* a) {
  return read_msg(r, wt, [&](int f, int w, Reader& rr) {
    switch (f) {
      case 1:
        a->symbol.clear();
        a->float_value = 0.f;
        a->string_value.clear();
        return parse_arg_value(rr, w, a);
      case 2:
        a->float_value = 0.f;
        a->string_value.clear();
        a->has_bools = false;
        return read_string(rr, w, &a->symbol);
      case 3:                   // func: ignored (ParseProtoArg never reads it)
        a->symbol.clear();
        a->float_value = 0.f;
        a->string_value.clear();
        return w == 2 && rr.skip(w);
      default: return rr.skip(w);
    }
  });
}

bool parse_operation(Reader& r, int wt, OperationPB* op) {
  return read_msg(r, wt, [&](int f, int w, Reader& rr) {
    switch (f) {
      case 1:
        return read_msg(rr, w, [&](int f2, int w2, Reader& r2) {
          return f2 == 1 ? read_string(r2, w2, &op->gate_id) : r2.skip(w2);
        });
      case 2: {
        ArgPB a;
        if (!read_msg(rr, w, [&](int f2, int w2, Reader& r2) {
              if (f2 == 1) return read_string(r2, w2, &a.key);
              if (f2 == 2) return parse_arg(r2, w2, &a);
              return r2.skip(w2);
            }))
          return false;
        op->args.push_back(std::move(a));
        return true;
      }
      case 3: {
        std::string id;
        if (!read_msg(rr, w, [&](int f2, int w2, Reader& r2) {
              return f2 == 2 ? read_string(r2, w2, &id) : r2.skip(w2);
            }))
          return false;
        op->qubits.push_back(std::move(id));
        return true;
      }
      default: return rr.skip(w);
    }
  });
}

bool parse_program_binary(const uint8_t* s, size_t n, ProgramPB* out) {
  return each_field(s, n, [&](int f, int w, Reader& r) {
    if (f == 2) {  // circuit
      return read_msg(r, w, [&](int f2, int w2, Reader& r2) {
        if (f2 == 1) { if (w2 != 0) return false; r2.varint(); return true; }
        if (f2 == 2) {
          MomentPB m;
          if (!read_msg(r2, w2, [&](int f3, int w3, Reader& r3) {
                if (f3 != 1) return r3.skip(w3);
                OperationPB op;
                if (!parse_operation(r3, w3, &op)) return false;
                m.operations.push_back(std::move(op));
                return true;
              }))
            return false;
          out->moments.push_back(std::move(m));
          return true;
        }
        return r2.skip(w2);
      });
    }
    if (f == 1 || f == 3) {
      // language / schedule: structurally validated (length-delimited) only.
      if (f == 3) out->moments.clear();  // oneof program: schedule wins
      return w == 2 && r.skip(w);
    }
    return r.skip(w);
  });
}

bool parse_pauli_sum_binary(const uint8_t* s, size_t n, PauliSumPB* out) {
  return each_field(s, n, [&](int f, int w, Reader& r) {
    if (f != 1) return r.skip(w);
    PauliTermPB t;
    if (!read_msg(r, w, [&](int f2, int w2, Reader& r2) {
          switch (f2) {
            case 1: if (w2 != 5) return false;
                    t.coefficient_real = r2.fixed32f(); return true;
            case 2: if (w2 != 5) return false; r2.fixed32f(); return true;
            case 3: {
              PauliPairPB p;
              if (!read_msg(r2, w2, [&](int f3, int w3, Reader& r3) {
                    if (f3 == 1) return read_string(r3, w3, &p.qubit_id);
                    if (f3 == 2) return read_string(r3, w3, &p.pauli_type);
                    return r3.skip(w3);
                  }))
                return false;
              t.paulis.push_back(std::move(p));
              return true;
            }
            default: return r2.skip(w2);
          }
        }))
      return false;
    out->terms.push_back(std::move(t));
    return true;
  });
}

// ------------------------------------------------------------- text format
// Generic tree: message = ordered list of (name, scalar | submessage).
struct TNode;
struct TField {
  std::string name;
  std::string scalar;                 // for scalar values (strings unescaped)
  std::unique_ptr<TNode> msg;         // for nested messages
};
struct TNode {
  std::vector<TField> fields;
};

struct TextLexer {
  const char* p;
  const char* end;
  void ws() {
    for (;;) {
      while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' ||
                         *p == ',' || *p == ';'))
        ++p;
      if (p < end && *p == '#') {
        while (p < end && *p != '\n') ++p;
        continue;
      }
      return;
    }
  }
  bool ident(std::string* s) {
    ws();
    const char* b = p;
    while (p < end && (isalnum((unsigned char)*p) || *p == '_' || *p == '.' ||
                       *p == '-' || *p == '+'))
      ++p;
    if (p == b) return false;
    s->assign(b, p - b);
    return true;
  }
  bool quoted(std::string* s) {
    ws();
    if (p >= end || (*p != '"' && *p != '\'')) return false;
    s->clear();
    // adjacent string literals concatenate
    while (p < end && (*p == '"' || *p == '\'')) {
      const char q = *p++;
      while (p < end && *p != q) {
        char c = *p++;
        if (c == '\n') return false;
        if (c == '\\') {
          if (p >= end) return false;
          c = *p++;
          switch (c) {
            case 'n': c = '\n'; break;
            case 't': c = '\t'; break;
            case 'r': c = '\r'; break;
            case '\\': case '\'': case '"': break;
            default:
              if (c >= '0' && c <= '7') {
                int v = c - '0';
                for (int k = 0; k < 2 && p < end && *p >= '0' && *p <= '7'; ++k)
                  v = v * 8 + (*p++ - '0');
                c = char(v);
              } else if (c == 'x') {
                int v = 0, k = 0;
                while (k < 2 && p < end && isxdigit((unsigned char)*p)) {
                  const char h = *p++;
                  v = v * 16 + (isdigit((unsigned char)h) ? h - '0'
                                : (tolower(h) - 'a' + 10));
                  ++k;
                }
                if (k == 0) return false;
                c = char(v);
              } else {
                return false;
              }
          }
        }
        s->push_back(c);
      }
      if (p >= end) return false;
      ++p;  // closing quote
      ws();
    }
    return true;
  }
};

bool parse_text_message(TextLexer& lx, TNode* node, char closer, int depth) {
  if (depth > 64) return false;
  for (;;) {
    lx.ws();
    if (lx.p >= lx.end) return closer == 0;
    if (closer && *lx.p == closer) { ++lx.p; return true; }
    TField f;
    if (!lx.ident(&f.name)) return false;
    lx.ws();
    bool colon = false;
    if (lx.p < lx.end && *lx.p == ':') { colon = true; ++lx.p; lx.ws(); }
    if (lx.p >= lx.end) return false;
    if (*lx.p == '{' || *lx.p == '<') {
      const char c = *lx.p == '{' ? '}' : '>';
      ++lx.p;
      f.msg.reset(new TNode);
      if (!parse_text_message(lx, f.msg.get(), c, depth + 1)) return false;
    } else {
      if (!colon) return false;
      if (*lx.p == '"' || *lx.p == '\'') {
        if (!lx.quoted(&f.scalar)) return false;
      } else if (!lx.ident(&f.scalar)) {
        return false;
      }
    }
    node->fields.push_back(std::move(f));
  }
}

bool text_float(const std::string& s, float* out) {
  if (s.empty()) return false;
  std::string t = s;
  if (t.back() == 'f' || t.back() == 'F') t.pop_back();
  if (t == "inf" || t == "infinity") { *out = INFINITY; return true; }
  if (t == "-inf" || t == "-infinity") { *out = -INFINITY; return true; }
  if (t == "nan") { *out = NAN; return true; }
  char* e = nullptr;
  const double d = std::strtod(t.c_str(), &e);
  if (e == t.c_str() || *e != 0) return false;
  *out = float(d);
  return true;
}

// schema-directed conversion; unknown field names are errors, as in
// TextFormat::ParseFromString.
bool text_arg_value(const TNode& n, ArgPB* a) {
  for (const auto& f : n.fields) {
    if (f.name == "float_value") {
      if (f.msg || !text_float(f.scalar, &a->float_value)) return false;
      a->string_value.clear();
    } else if (f.name == "string_value") {
      if (f.msg) return false;
      a->string_value = f.scalar;
      a->float_value = 0.f;
    } else if (f.name == "double_value") {
      float unused;
      if (f.msg || !text_float(f.scalar, &unused)) return false;
    } else if (f.name == "bool_values") {
      if (!f.msg) return false;
      // RepeatedBoolean{repeated bool values = 1} as proto3 writes it (packed)
      std::string packed;
      for (const auto& v : f.msg->fields) {
        if (v.name != "values" || v.msg) return false;
        if (v.scalar == "true" || v.scalar == "1") packed.push_back(char(1));
        else if (v.scalar == "false" || v.scalar == "0") packed.push_back(char(0));
        else return false;
      }
      a->has_bools = true;
      a->bools_wire.clear();
      if (!packed.empty()) {
        a->bools_wire.push_back(char(0x0A));              // field 1, length-delimited
        a->bools_wire.push_back(char(packed.size()));     // < 128 values
        a->bools_wire += packed;
      }
      a->float_value = 0.f;
      a->string_value.clear();
    } else {
      return false;
    }
  }
  return true;
}

bool text_arg(const TNode& n, ArgPB* a) {
  for (const auto& f : n.fields) {
    if (f.name == "arg_value") {
      if (!f.msg || !text_arg_value(*f.msg, a)) return false;
      a->symbol.clear();
    } else if (f.name == "symbol") {
      if (f.msg) return false;
      a->symbol = f.scalar;
    } else if (f.name == "func") {
      if (!f.msg) return false;
    } else {
      return false;
    }
  }
  return true;
}

bool text_operation(const TNode& n, OperationPB* op) {
  for (const auto& f : n.fields) {
    if (!f.msg) return false;
    if (f.name == "gate") {
      for (const auto& g : f.msg->fields) {
        if (g.name != "id" || g.msg) return false;
        op->gate_id = g.scalar;
      }
    } else if (f.name == "args") {
      ArgPB a;
      for (const auto& e : f.msg->fields) {
        if (e.name == "key" && !e.msg) a.key = e.scalar;
        else if (e.name == "value" && e.msg) { if (!text_arg(*e.msg, &a)) return false; }
        else return false;
      }
      op->args.push_back(std::move(a));
    } else if (f.name == "qubits") {
      std::string id;
      for (const auto& g : f.msg->fields) {
        if (g.name != "id" || g.msg) return false;
        id = g.scalar;
      }
      op->qubits.push_back(std::move(id));
    } else {
      return false;
    }
  }
  return true;
}

bool parse_program_text(const char* s, size_t n, ProgramPB* out) {
  TextLexer lx{s, s + n};
  TNode root;
  if (!parse_text_message(lx, &root, 0, 0)) return false;
  for (const auto& f : root.fields) {
    if (f.name == "language" || f.name == "schedule") {
      if (!f.msg) return false;
    } else if (f.name == "circuit") {
      if (!f.msg) return false;
      for (const auto& c : f.msg->fields) {
        if (c.name == "scheduling_strategy") {
          if (c.msg) return false;
        } else if (c.name == "moments") {
          if (!c.msg) return false;
          MomentPB m;
          for (const auto& o : c.msg->fields) {
            if (o.name != "operations" || !o.msg) return false;
            OperationPB op;
            if (!text_operation(*o.msg, &op)) return false;
            m.operations.push_back(std::move(op));
          }
          out->moments.push_back(std::move(m));
        } else {
          return false;
        }
      }
    } else {
      return false;
    }
  }
  return true;
}

bool parse_pauli_sum_text(const char* s, size_t n, PauliSumPB* out) {
  TextLexer lx{s, s + n};
  TNode root;
  if (!parse_text_message(lx, &root, 0, 0)) return false;
  for (const auto& f : root.fields) {
    if (f.name != "terms" || !f.msg) return false;
    PauliTermPB t;
    for (const auto& c : f.msg->fields) {
      if (c.name == "coefficient_real") {
        if (c.msg || !text_float(c.scalar, &t.coefficient_real)) return false;
      } else if (c.name == "coefficient_imag") {
        float unused;
        if (c.msg || !text_float(c.scalar, &unused)) return false;
      } else if (c.name == "paulis") {
        if (!c.msg) return false;
        PauliPairPB p;
        for (const auto& q : c.msg->fields) {
          if (q.msg) return false;
          if (q.name == "qubit_id") p.qubit_id = q.scalar;
          else if (q.name == "pauli_type") p.pauli_type = q.scalar;
          else return false;
        }
        t.paulis.push_back(std::move(p));
      } else {
        return false;
      }
    }
    out->terms.push_back(std::move(t));
  }
  return true;
}

}  // namespace

bool ParseProgram(const char* data, size_t len, ProgramPB* out) {
  out->moments.clear();
  if (parse_program_binary(reinterpret_cast<const uint8_t*>(data), len, out))
    return true;
  out->moments.clear();
  if (parse_program_text(data, len, out)) return true;
  out->moments.clear();
  return false;
}

bool ParsePauliSum(const char* data, size_t len, PauliSumPB* out) {
  out->terms.clear();
  if (parse_pauli_sum_binary(reinterpret_cast<const uint8_t*>(data), len, out))
    return true;
  out->terms.clear();
  if (parse_pauli_sum_text(data, len, out)) return true;
  out->terms.clear();
  return false;
}

}  // namespace tfqb
