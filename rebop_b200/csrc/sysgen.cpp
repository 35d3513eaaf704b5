// sysgen.cpp -- the `define_system!` DSL (src/gillespie_macro.rs:49-61) on the host.
//
//   params...;
//   Name { species, ... }
//   rname : [n] A + [n] B => [n] C + ... @ rate
//
// The reference expands this at the user's compile time into a struct with one field per species
// and per parameter and a fully unrolled advance_until (src/gillespie_macro.rs:62-126).  Here the
// same text is parsed into a `rebop_system`; `rebop_system_network` plays `Name::with_parameters`
// (it evaluates every rate expression over the parameter values and lowers the reactions in
// define_system! arithmetic), and the build-time tool `rebop_sysgen` (sysgen_main.cpp) turns the
// text into the CUDA source of a network-specialised kernel that nvcc compiles into the library --
// the stand-in for a `build.rs` step.
#include "sysgen.hpp"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>

namespace {

struct Lexer {
  const char* s;
  size_t pos = 0;
  std::string error;

  explicit Lexer(const char* text) : s(text) {}
  static bool ident_start(unsigned char c) { return std::isalpha(c) || c == '_' || c >= 0x80; }
  static bool ident_char(unsigned char c) { return std::isalnum(c) || c == '_' || c >= 0x80; }
  void skip() {
    for (;;) {
      while (s[pos] && std::isspace((unsigned char)s[pos])) ++pos;
      if (s[pos] == '/' && s[pos + 1] == '/') {  // Rust line comment
        while (s[pos] && s[pos] != '\n') ++pos;
        continue;
      }
      break;
    }
  }
  bool eof() { skip(); return s[pos] == 0; }
  bool peek(const char* tok) { skip(); return std::strncmp(s + pos, tok, std::strlen(tok)) == 0; }
  bool accept(const char* tok) {
    if (!peek(tok)) return false;
    pos += std::strlen(tok);
    return true;
  }
  bool peek_ident() { skip(); return ident_start((unsigned char)s[pos]); }
  std::string ident() {
    skip();
    size_t b = pos;
    if (!ident_start((unsigned char)s[pos])) return std::string();
    while (ident_char((unsigned char)s[pos])) ++pos;
    return std::string(s + b, pos - b);
  }
  bool peek_number() { skip(); return std::isdigit((unsigned char)s[pos]) || (s[pos] == '.' && std::isdigit((unsigned char)s[pos + 1])); }
  // Rust float/integer literal: digits with optional '_' separators, fraction, exponent, f64 suffix
  bool number(double* out, bool* is_int) {
    skip();
    std::string buf;
    size_t p = pos;
    bool integer = true;
    auto digits = [&] {
      while (std::isdigit((unsigned char)s[p]) || s[p] == '_') {
        if (s[p] != '_') buf.push_back(s[p]);
        ++p;
      }
    };
    digits();
    if (s[p] == '.' && !ident_start((unsigned char)s[p + 1])) {
      integer = false;
      buf.push_back('.');
      ++p;
      digits();
    }
    if ((s[p] == 'e' || s[p] == 'E') && (std::isdigit((unsigned char)s[p + 1]) ||
                                          ((s[p + 1] == '+' || s[p + 1] == '-') && std::isdigit((unsigned char)s[p + 2])))) {
      integer = false;
      buf.push_back('e');
      ++p;
      if (s[p] == '+' || s[p] == '-') buf.push_back(s[p++]);
      digits();
    }
    if (std::strncmp(s + p, "f64", 3) == 0 || std::strncmp(s + p, "_f64", 4) == 0) {
      integer = false;
      p += s[p] == '_' ? 4 : 3;
    }
    if (buf.empty() || buf == ".") return false;
    *out = std::strtod(buf.c_str(), nullptr);
    if (is_int) *is_int = integer;
    pos = p;
    return true;
  }
};

// rate expression: + - * / over parameters and literals, unary minus, parentheses
std::unique_ptr<RbRateExpr> parse_expr(Lexer& lx, const rebop_system& sys);

std::unique_ptr<RbRateExpr> parse_atom(Lexer& lx, const rebop_system& sys) {
  auto node = std::make_unique<RbRateExpr>();
  if (lx.accept("(")) {
    node = parse_expr(lx, sys);
    if (node && !lx.accept(")")) {
      lx.error = "expected ')' in a rate expression";
      return nullptr;
    }
    return node;
  }
  if (lx.accept("-")) {
    node->kind = RbRateExpr::NEG;
    node->a = parse_atom(lx, sys);
    return node->a ? std::move(node) : nullptr;
  }
  if (lx.peek_number()) {
    node->kind = RbRateExpr::CONST;
    if (!lx.number(&node->value, nullptr)) {
      lx.error = "malformed number in a rate expression";
      return nullptr;
    }
    return node;
  }
  if (lx.peek_ident()) {
    const std::string name = lx.ident();
    for (size_t i = 0; i < sys.params.size(); ++i)
      if (sys.params[i] == name) {
        node->kind = RbRateExpr::PARAM;
        node->index = (int)i;
        return node;
      }
    for (const std::string& sp : sys.species)
      if (sp == name) {
        // The macro would read a snapshot of the species taken once per advance_until call
        // (src/gillespie_macro.rs:101-104); that stale-value quirk is not reproduced.
        lx.error = "rate expression of a define_system! reaction names species '" + name +
                   "'; only parameters and literals are supported";
        return nullptr;
      }
    lx.error = "unknown identifier '" + name + "' in a rate expression";
    return nullptr;
  }
  lx.error = "expected a rate expression";
  return nullptr;
}

std::unique_ptr<RbRateExpr> parse_term(Lexer& lx, const rebop_system& sys) {
  auto lhs = parse_atom(lx, sys);
  while (lhs) {
    int kind;
    if (lx.peek("*")) kind = RbRateExpr::MUL;
    else if (lx.peek("/") && !lx.peek("//")) kind = RbRateExpr::DIV;
    else break;
    lx.pos += 1;
    auto node = std::make_unique<RbRateExpr>();
    node->kind = kind;
    node->a = std::move(lhs);
    node->b = parse_atom(lx, sys);
    if (!node->b) return nullptr;
    lhs = std::move(node);
  }
  return lhs;
}

std::unique_ptr<RbRateExpr> parse_expr(Lexer& lx, const rebop_system& sys) {
  auto lhs = parse_term(lx, sys);
  while (lhs) {
    int kind;
    if (lx.peek("+")) kind = RbRateExpr::ADD;
    else if (lx.peek("-")) kind = RbRateExpr::SUB;
    else break;
    lx.pos += 1;
    auto node = std::make_unique<RbRateExpr>();
    node->kind = kind;
    node->a = std::move(lhs);
    node->b = parse_term(lx, sys);
    if (!node->b) return nullptr;
    lhs = std::move(node);
  }
  return lhs;
}

double eval_expr(const RbRateExpr& e, const double* params) {
  switch (e.kind) {
    case RbRateExpr::CONST: return e.value;
    case RbRateExpr::PARAM: return params[e.index];
    case RbRateExpr::NEG: return -eval_expr(*e.a, params);
    case RbRateExpr::ADD: return eval_expr(*e.a, params) + eval_expr(*e.b, params);
    case RbRateExpr::SUB: return eval_expr(*e.a, params) - eval_expr(*e.b, params);
    case RbRateExpr::MUL: return eval_expr(*e.a, params) * eval_expr(*e.b, params);
    case RbRateExpr::DIV: return eval_expr(*e.a, params) / eval_expr(*e.b, params);
  }
  return std::numeric_limits<double>::quiet_NaN();
}

// [n] A + [n] B ...   (possibly empty; ends at "=>" or "@")
bool parse_side(Lexer& lx, const rebop_system& sys, std::vector<std::pair<uint32_t, uint32_t>>* side) {
  if (lx.peek("=>") || lx.peek("@")) return true;
  for (;;) {
    uint32_t n = 1;
    if (lx.peek_number()) {
      double v;
      bool is_int = false;
      if (!lx.number(&v, &is_int) || !is_int || v < 0 || v > 255) {
        lx.error = "stoichiometric coefficient must be an integer literal in [0, 255]";
        return false;
      }
      n = (uint32_t)v;
    }
    const std::string name = lx.ident();
    if (name.empty()) {
      lx.error = "expected a species name";
      return false;
    }
    size_t idx = sys.species.size();
    for (size_t i = 0; i < sys.species.size(); ++i)
      if (sys.species[i] == name) idx = i;
    if (idx == sys.species.size()) {
      lx.error = "no field `" + name + "` on type `" + sys.name + "`";  // what rustc would say
      return false;
    }
    side->push_back({(uint32_t)idx, n});
    if (!lx.accept("+")) return true;
  }
}

int copy_string(const std::string& v, char* buf, size_t cap, size_t* needed) {
  if (needed) *needed = v.size() + 1;
  if (buf && cap) {
    const size_t n = v.size() + 1 < cap ? v.size() + 1 : cap;
    std::memcpy(buf, v.c_str(), n);
    buf[cap - 1 < n ? cap - 1 : n - 1] = 0;
  }
  return REBOP_OK;
}

}  // namespace

int rb_system_parse(const char* text, rebop_system* sys) {
  Lexer lx(text);
  while (!lx.peek(";")) {
    const std::string p = lx.ident();
    if (p.empty()) return rb_fail(REBOP_ERR_PARSE, "define_system: expected parameter names followed by ';'");
    sys->params.push_back(p);
  }
  lx.accept(";");
  sys->name = lx.ident();
  if (sys->name.empty() || !lx.accept("{")) return rb_fail(REBOP_ERR_PARSE, "define_system: expected `Name { species, ... }`");
  while (!lx.peek("}")) {
    const std::string sp = lx.ident();
    if (sp.empty()) return rb_fail(REBOP_ERR_PARSE, "define_system: expected a species name");
    sys->species.push_back(sp);
    if (!lx.accept(",") && !lx.peek("}")) return rb_fail(REBOP_ERR_PARSE, "define_system: expected ',' or '}' in the species list");
  }
  lx.accept("}");
  while (!lx.eof()) {
    RbSystemReaction rx;
    rx.name = lx.ident();
    if (rx.name.empty() || !lx.accept(":")) return rb_fail(REBOP_ERR_PARSE, "define_system: expected `reaction_name :`");
    if (!parse_side(lx, *sys, &rx.lhs)) return rb_fail(REBOP_ERR_PARSE, "define_system: " + lx.error);
    if (!lx.accept("=>")) return rb_fail(REBOP_ERR_PARSE, "define_system: expected `=>` in reaction " + rx.name);
    if (!parse_side(lx, *sys, &rx.rhs)) return rb_fail(REBOP_ERR_PARSE, "define_system: " + lx.error);
    if (!lx.accept("@")) return rb_fail(REBOP_ERR_PARSE, "define_system: expected `@ rate` in reaction " + rx.name);
    rx.rate = parse_expr(lx, *sys);
    if (!rx.rate) return rb_fail(REBOP_ERR_PARSE, "define_system: reaction " + rx.name + ": " + lx.error);
    sys->reactions.push_back(std::move(rx));
  }
  return REBOP_OK;
}

// Name::with_parameters(...) (src/gillespie_macro.rs:86-95) + the reaction lowering the macro
// performs syntactically: rate * _rate_lma!(n * r) * ... in the order written (:106), and
// `self.r -= n`, `self.p += n` (:160-163).
int rb_system_network(const rebop_system& sys, const double* params, size_t n_params, rebop_network* net) {
  if (n_params != sys.params.size())
    return rb_fail(REBOP_ERR_INVALID, sys.name + "::with_parameters takes " + std::to_string(sys.params.size()) +
                                          " parameters, " + std::to_string(n_params) + " given");
  net->n_species = (uint32_t)sys.species.size();
  net->arith = REBOP_ARITH_MACRO;
  net->rx.clear();
  for (const RbSystemReaction& r : sys.reactions) {
    RbReaction rx;
    rx.k = eval_expr(*r.rate, params);
    rx.diff.assign(sys.species.size(), 0);
    // `$self.$r -= 1 $(+ $nr - 1)?`: an explicit coefficient n moves the count by n (so `0 A` by 0),
    // no coefficient by 1 (stored as n = 1); _rate_lma!(0 * A) is A (src/gillespie_macro.rs:133-146)
    for (const auto& term : r.lhs) {
      rx.term_idx.push_back(term.first);
      rx.term_exp.push_back(term.second);
      rx.diff[term.first] -= (int64_t)term.second;
    }
    for (const auto& term : r.rhs) rx.diff[term.first] += (int64_t)term.second;
    net->rx.push_back(std::move(rx));
  }
  return REBOP_OK;
}

// ---- C ABI ----
extern "C" int rebop_system_parse(const char* dsl_text, rebop_system** out) {
  if (!dsl_text || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::unique_ptr<rebop_system> sys(new rebop_system());
  int st = rb_system_parse(dsl_text, sys.get());
  if (st) return st;
  *out = sys.release();
  return REBOP_OK;
}
extern "C" void rebop_system_destroy(rebop_system* sys) { delete sys; }
extern "C" int rebop_system_name(const rebop_system* sys, char* buf, size_t cap, size_t* needed) {
  if (!sys) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  return copy_string(sys->name, buf, cap, needed);
}
extern "C" int rebop_system_counts(const rebop_system* sys, uint32_t* n_params, uint32_t* n_species,
                                   uint32_t* n_reactions) {
  if (!sys) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (n_params) *n_params = (uint32_t)sys->params.size();
  if (n_species) *n_species = (uint32_t)sys->species.size();
  if (n_reactions) *n_reactions = (uint32_t)sys->reactions.size();
  return REBOP_OK;
}
extern "C" int rebop_system_param_name(const rebop_system* sys, uint32_t i, char* buf, size_t cap, size_t* needed) {
  if (!sys || i >= sys->params.size()) return rb_fail(REBOP_ERR_OUT_OF_RANGE, "parameter index out of range");
  return copy_string(sys->params[i], buf, cap, needed);
}
extern "C" int rebop_system_species_name(const rebop_system* sys, uint32_t i, char* buf, size_t cap, size_t* needed) {
  if (!sys || i >= sys->species.size()) return rb_fail(REBOP_ERR_OUT_OF_RANGE, "species index out of range");
  return copy_string(sys->species[i], buf, cap, needed);
}
extern "C" int rebop_system_reaction_name(const rebop_system* sys, uint32_t i, char* buf, size_t cap, size_t* needed) {
  if (!sys || i >= sys->reactions.size()) return rb_fail(REBOP_ERR_OUT_OF_RANGE, "reaction index out of range");
  return copy_string(sys->reactions[i].name, buf, cap, needed);
}
extern "C" int rebop_system_rates(const rebop_system* sys, const double* params, size_t n_params, double* rates) {
  if (!sys || (!params && n_params) || (!rates && !sys->reactions.empty())) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (n_params != sys->params.size()) return rb_fail(REBOP_ERR_INVALID, "wrong number of parameters");
  for (size_t r = 0; r < sys->reactions.size(); ++r) rates[r] = eval_expr(*sys->reactions[r].rate, params);
  return REBOP_OK;
}
extern "C" int rebop_system_network(const rebop_system* sys, const double* params, size_t n_params, rebop_network** out) {
  if (!sys || !out || (!params && n_params)) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::unique_ptr<rebop_network> net(new rebop_network());
  int st = rb_system_network(*sys, params, n_params, net.get());
  if (st) return st;
  *out = net.release();
  return REBOP_OK;
}
