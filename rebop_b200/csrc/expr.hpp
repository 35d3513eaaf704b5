// expr.hpp -- rate-expression front end (the reference's PExpr, src/expr.rs:43-273).
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "network.hpp"

struct RbPExpr {
  enum Kind { Constant, Variable, Neg, Add, Sub, Mul, Div, Pow, Max, Min, Exp } kind = Constant;
  double value = 0.0;
  std::string name;
  std::unique_ptr<RbPExpr> a, b;
};

// "...".parse::<PExpr>(): nullptr when the text is not understood.
std::unique_ptr<RbPExpr> rb_pexpr_parse(const std::string& text);
// Display for PExpr.
std::string rb_pexpr_format(const RbPExpr& e);
// PExpr::to_expr, emitted as a post-order program appended to *out.
int rb_pexpr_lower(const RbPExpr& e, const std::vector<std::string>& species,
                   const std::vector<std::pair<std::string, double>>& params, std::vector<rebop_expr_op>* out);
