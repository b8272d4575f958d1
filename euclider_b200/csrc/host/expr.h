// Infix expression -> RPN compiler for ComponentTransformationExpr strings
// (reference: `meval::Expr::from_str`, src/scene.rs:964-981; evaluated per ray transition in
// material.rs:91-112).  meval 0.1.0 is not vendored: grammar and precedences are RECOLLECTION
// (+,- < *,/,% < unary < ^ (right assoc); constants pi, e; usual one-argument functions).
#pragma once
#include <string>
#include <vector>

#include "euclider_b200.h"

namespace eucl {

// `legend` maps variable names (single characters, material.rs:76-88) to vector components.
// Returns false and sets *error for syntax errors and unknown identifiers.
bool expr_compile(const std::string& text, const std::string& legend, int dim,
                  std::vector<EuclExprOp>* out, std::string* error);

// Host-side evaluation (used for load-time checks and unit tests only).
double expr_eval(const EuclExprOp* ops, int len, const double* vars);

} // namespace eucl
