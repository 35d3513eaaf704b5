// expr.cpp -- rate-expression front end: parse, print, lower.
//
// Restates the reference's `PExpr` (src/expr.rs:43-141) and its winnow grammar
// (src/expr.rs:144-273) as a hand-written backtracking recursive-descent parser with the same
// alternatives in the same order:
//
//   expr    = term   (space0 [+-] space0 term)*          add_sub, left fold        :205-217
//   term    = factor (space0 [*/] space0 factor)*        mul_div, left fold        :219-231
//   factor  = atom space0 '^' space0 atom | atom         pow is NOT chained        :233-237,179-181
//   atom    = exp | max | min | variable | constant | '(' space0 expr space0 ')' | neg   :183-187
//   neg     = '-' space0 term                            unary minus binds a whole term  :189-193
//   variable= [A-Za-z_][A-Za-z0-9_]*   (before constant, so `inf`, `nan`, `e` are names)  :157-169
//   constant= winnow `float`           (sign, digits[.digits] | .digits, exponent; nan/inf/infinity)
//
// The whole input must be consumed (`expr.parse(s)`, :270-272); there is no leading or trailing
// white space at top level.  A separator followed by something that is not an operand is not
// consumed (winnow's separated_foldl1 backtracks over it), so "1+" fails at end-of-input.
#include "expr.hpp"

#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

struct Parser {
  const char* s;
  size_t n;
  size_t pos = 0;
  bool fatal = false;  // winnow `cut_err`: an error that alternatives must not recover from

  bool eof() const { return pos >= n; }
  char peek() const { return pos < n ? s[pos] : '\0'; }
  void space0() {
    while (pos < n && (s[pos] == ' ' || s[pos] == '\t')) ++pos;
  }
  bool lit(const char* word) {
    size_t len = std::strlen(word);
    if (n - pos >= len && std::memcmp(s + pos, word, len) == 0) {
      pos += len;
      return true;
    }
    return false;
  }
  static bool is_alpha(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
  static bool is_digit(char c) { return c >= '0' && c <= '9'; }

  typedef std::unique_ptr<RbPExpr> Node;

  static Node make(RbPExpr::Kind k, Node a = nullptr, Node b = nullptr) {
    Node e(new RbPExpr());
    e->kind = k;
    e->a = std::move(a);
    e->b = std::move(b);
    return e;
  }

  Node expr() {
    Node left = term();
    if (!left) return nullptr;
    for (;;) {
      size_t save = pos;
      space0();
      char op = peek();
      if (op != '+' && op != '-') { pos = save; return left; }
      ++pos;
      space0();
      Node right = term();
      if (fatal) return nullptr;
      if (!right) { pos = save; return left; }
      left = make(op == '+' ? RbPExpr::Add : RbPExpr::Sub, std::move(left), std::move(right));
    }
  }

  Node term() {
    Node left = factor();
    if (!left) return nullptr;
    for (;;) {
      size_t save = pos;
      space0();
      char op = peek();
      if (op != '*' && op != '/') { pos = save; return left; }
      ++pos;
      space0();
      Node right = factor();
      if (fatal) return nullptr;
      if (!right) { pos = save; return left; }
      left = make(op == '*' ? RbPExpr::Mul : RbPExpr::Div, std::move(left), std::move(right));
    }
  }

  Node factor() {
    size_t save = pos;
    Node base = atom();
    if (!base) return nullptr;
    size_t after_atom = pos;
    space0();
    if (peek() == '^') {
      ++pos;
      space0();
      Node exponent = atom();
      if (fatal) return nullptr;
      if (exponent) return make(RbPExpr::Pow, std::move(base), std::move(exponent));
    }
    // pow failed: alt() retries `atom` from the start, which yields the same node
    (void)save;
    pos = after_atom;
    return base;
  }

  Node parentheses() {
    size_t save = pos;
    if (peek() != '(') return nullptr;
    ++pos;
    space0();
    Node inner = expr();
    if (inner) {
      space0();
      if (peek() == ')') {
        ++pos;
        return inner;
      }
    }
    pos = save;
    return nullptr;
  }

  Node call1(const char* name, RbPExpr::Kind kind) {
    size_t save = pos;
    if (lit(name)) {
      Node arg = parentheses();
      if (arg) return make(kind, std::move(arg));
    }
    pos = save;
    return nullptr;
  }

  Node call2(const char* name, RbPExpr::Kind kind) {
    size_t save = pos;
    if (lit(name) && peek() == '(') {
      ++pos;
      space0();
      Node a = expr();
      if (a) {
        space0();
        if (peek() == ',') {
          ++pos;
          space0();
          Node b = expr();
          if (b) {
            space0();
            if (peek() == ')') {
              ++pos;
              return make(kind, std::move(a), std::move(b));
            }
          }
        }
      }
    }
    pos = save;
    return nullptr;
  }

  Node variable() {
    if (!(is_alpha(peek()) || peek() == '_')) return nullptr;
    size_t start = pos++;
    while (pos < n && (is_alpha(s[pos]) || is_digit(s[pos]) || s[pos] == '_')) ++pos;
    Node e = make(RbPExpr::Variable);
    e->name.assign(s + start, pos - start);
    return e;
  }

  static bool caseless(const char* p, size_t avail, const char* word) {
    size_t len = std::strlen(word);
    if (avail < len) return false;
    for (size_t i = 0; i < len; ++i) {
      char c = p[i];
      if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
      if (c != word[i]) return false;
    }
    return true;
  }

  // winnow::ascii::float
  Node constant() {
    size_t start = pos, q = pos;
    if (q < n && (s[q] == '+' || s[q] == '-')) ++q;
    size_t digits_start = q;
    bool number = false;
    if (q < n && is_digit(s[q])) {
      while (q < n && is_digit(s[q])) ++q;
      if (q < n && s[q] == '.') {
        ++q;
        while (q < n && is_digit(s[q])) ++q;
      }
      number = true;
    } else if (q + 1 < n && s[q] == '.' && is_digit(s[q + 1])) {
      ++q;
      while (q < n && is_digit(s[q])) ++q;
      number = true;
    }
    if (number) {
      if (q < n && (s[q] == 'e' || s[q] == 'E')) {
        size_t r = q + 1;
        if (r < n && (s[r] == '+' || s[r] == '-')) ++r;
        if (!(r < n && is_digit(s[r]))) {  // cut_err(digit1)
          fatal = true;
          return nullptr;
        }
        while (r < n && is_digit(s[r])) ++r;
        q = r;
      }
    } else {
      // exceptions: nan (unsigned), [+-]infinity, [+-]inf -- case-insensitive
      if (caseless(s + start, n - start, "nan")) q = start + 3;
      else if (caseless(s + digits_start, n - digits_start, "infinity")) q = digits_start + 8;
      else if (caseless(s + digits_start, n - digits_start, "inf")) q = digits_start + 3;
      else return nullptr;
    }
    std::string text(s + start, q - start);
    Node e = make(RbPExpr::Constant);
    e->value = std::strtod(text.c_str(), nullptr);  // correctly rounded, like str::parse::<f64>
    pos = q;
    return e;
  }

  Node neg() {
    size_t save = pos;
    if (peek() != '-') return nullptr;
    ++pos;
    space0();
    Node inner = term();
    if (inner) return make(RbPExpr::Neg, std::move(inner));
    pos = save;
    return nullptr;
  }

  Node atom() {
    Node e;
    if ((e = call1("exp", RbPExpr::Exp))) return e;
    if (fatal) return nullptr;
    if ((e = call2("max", RbPExpr::Max))) return e;
    if (fatal) return nullptr;
    if ((e = call2("min", RbPExpr::Min))) return e;
    if (fatal) return nullptr;
    if ((e = variable())) return e;
    if ((e = constant())) return e;
    if (fatal) return nullptr;
    if ((e = parentheses())) return e;
    if (fatal) return nullptr;
    return neg();
  }
};

// Rust's `{}` for f64: shortest digits that round-trip, never scientific notation.
std::string format_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  char buf[512];
  auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  return std::string(buf, res.ptr);
}

}  // namespace

std::unique_ptr<RbPExpr> rb_pexpr_parse(const std::string& text) {
  Parser p{text.data(), text.size()};
  std::unique_ptr<RbPExpr> e = p.expr();
  if (!e || p.fatal || !p.eof()) return nullptr;
  return e;
}

std::string rb_pexpr_format(const RbPExpr& e) {
  switch (e.kind) {
    case RbPExpr::Constant: return format_f64(e.value);
    case RbPExpr::Variable: return e.name;
    case RbPExpr::Neg: return "(-" + rb_pexpr_format(*e.a) + ")";
    case RbPExpr::Add: return "(" + rb_pexpr_format(*e.a) + " + " + rb_pexpr_format(*e.b) + ")";
    case RbPExpr::Sub: return "(" + rb_pexpr_format(*e.a) + " - " + rb_pexpr_format(*e.b) + ")";
    case RbPExpr::Mul: return "(" + rb_pexpr_format(*e.a) + " * " + rb_pexpr_format(*e.b) + ")";
    case RbPExpr::Div: return "(" + rb_pexpr_format(*e.a) + " / " + rb_pexpr_format(*e.b) + ")";
    case RbPExpr::Pow: return "(" + rb_pexpr_format(*e.a) + " ^ " + rb_pexpr_format(*e.b) + ")";
    case RbPExpr::Max: return "max(" + rb_pexpr_format(*e.a) + ", " + rb_pexpr_format(*e.b) + ")";
    case RbPExpr::Min: return "min(" + rb_pexpr_format(*e.a) + ", " + rb_pexpr_format(*e.b) + ")";
    case RbPExpr::Exp: return "exp(" + rb_pexpr_format(*e.a) + ")";
  }
  return "";
}

// PExpr::to_expr (src/expr.rs:58-109), emitted directly as the post-order program.
int rb_pexpr_lower(const RbPExpr& e, const std::vector<std::string>& species,
                   const std::vector<std::pair<std::string, double>>& params, std::vector<rebop_expr_op>* out) {
  rebop_expr_op op;
  op.index = 0;
  op.value = 0.0;
  switch (e.kind) {
    case RbPExpr::Constant:
      op.op = REBOP_OP_CONST;
      op.value = e.value;
      out->push_back(op);
      return REBOP_OK;
    case RbPExpr::Variable: {
      for (size_t i = 0; i < species.size(); ++i)
        if (species[i] == e.name) {  // species shadow parameters (:66-67)
          op.op = REBOP_OP_SPECIES;
          op.index = (int32_t)i;
          out->push_back(op);
          return REBOP_OK;
        }
      for (const auto& kv : params)
        if (kv.first == e.name) {
          op.op = REBOP_OP_CONST;
          op.value = kv.second;
          out->push_back(op);
          return REBOP_OK;
        }
      return rb_fail(REBOP_ERR_MISSING_PARAM, "Parameter " + e.name + " should have a value");
    }
    default: break;
  }
  int st = rb_pexpr_lower(*e.a, species, params, out);
  if (st) return st;
  if (e.b) {
    st = rb_pexpr_lower(*e.b, species, params, out);
    if (st) return st;
  }
  switch (e.kind) {
    case RbPExpr::Neg: op.op = REBOP_OP_NEG; break;
    case RbPExpr::Add: op.op = REBOP_OP_ADD; break;
    case RbPExpr::Sub: op.op = REBOP_OP_SUB; break;
    case RbPExpr::Mul: op.op = REBOP_OP_MUL; break;
    case RbPExpr::Div: op.op = REBOP_OP_DIV; break;
    case RbPExpr::Pow: op.op = REBOP_OP_POW; break;
    case RbPExpr::Max: op.op = REBOP_OP_MAX; break;
    case RbPExpr::Min: op.op = REBOP_OP_MIN; break;
    case RbPExpr::Exp: op.op = REBOP_OP_EXP; break;
    default: break;
  }
  out->push_back(op);
  return REBOP_OK;
}

// ---- C ABI ----
struct rebop_pexpr {
  std::unique_ptr<RbPExpr> root;
};

extern "C" int rebop_pexpr_parse(const char* text, rebop_pexpr** out) {
  if (!text || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::unique_ptr<RbPExpr> e = rb_pexpr_parse(text);
  if (!e) return rb_fail(REBOP_ERR_PARSE, "Rate expression not understood");
  *out = new rebop_pexpr{std::move(e)};
  return REBOP_OK;
}

extern "C" void rebop_pexpr_destroy(rebop_pexpr* e) { delete e; }

extern "C" int rebop_pexpr_format(const rebop_pexpr* e, char* buf, size_t cap, size_t* needed) {
  if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  const std::string text = rb_pexpr_format(*e->root);
  if (needed) *needed = text.size() + 1;
  if (buf && cap) {
    const size_t ncopy = text.size() < cap - 1 ? text.size() : cap - 1;
    std::memcpy(buf, text.data(), ncopy);
    buf[ncopy] = '\0';
  }
  return REBOP_OK;
}

extern "C" int rebop_pexpr_lower(const rebop_pexpr* e, const char* const* species_names, size_t n_species,
                                 const char* const* param_names, const double* param_values, size_t n_params,
                                 rebop_expr_op* program, size_t cap, size_t* n_ops) {
  if (!e || !n_ops) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::vector<std::string> species(n_species);
  for (size_t i = 0; i < n_species; ++i) species[i] = species_names[i];
  std::vector<std::pair<std::string, double>> params(n_params);
  for (size_t i = 0; i < n_params; ++i) params[i] = {param_names[i], param_values[i]};
  std::vector<rebop_expr_op> prog;
  int st = rb_pexpr_lower(*e->root, species, params, &prog);
  if (st) return st;
  *n_ops = prog.size();
  if (program && cap >= prog.size()) std::memcpy(program, prog.data(), prog.size() * sizeof(rebop_expr_op));
  return REBOP_OK;
}
