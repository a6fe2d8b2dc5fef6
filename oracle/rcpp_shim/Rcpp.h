// Minimal stand-in for <Rcpp.h> -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Purpose: compile the reference's own two kernel sources, UNMODIFIED and where they lie
//   /root/reference/src/covariance_functionsC.cpp
//   /root/reference/src/covariance_function_derivativesC.cpp
// into oracle/_ref/libsparseRGPs_ref.so without an R / Rcpp installation (neither exists in this image,
// SURVEY.md 8c), so that the C restatement in oracle/ref_kernels.c and the CUDA kernels can be checked against
// the arithmetic the reference itself executes. Only the subset of the Rcpp API those two files touch is
// provided, with Rcpp's semantics where they can influence a result:
//   * NumericVector / NumericMatrix share storage on copy (SEXP semantics), matrices are column-major,
//     constructors zero-fill;
//   * sugar expressions (x1 - x2, pow(v, 2), exp(v), sum(v), abs(v), v == w, all(), any(), is_true()) evaluate
//     element by element in double, `sum` accumulates left to right in double (Rcpp::sugar::Sum), `pow` calls
//     ::pow(double, double) (Rcpp::sugar::Pow), comparisons involving NaN/NA give NA and all()/any() follow
//     R's three-valued logic;
//   * List is a named list looked up BY NAME (index_out_of_bounds when absent), as<double> insists on length 1;
//   * Function("name") resolves R-level closures from a registry the glue fills with the R wrappers the package
//     itself would bind (R/RcppExports.R: real_to_pos, real_to_bounded -> the same translation unit's C++).
// What this header is NOT: Rcpp. Expression templates are evaluated eagerly (same per-element operation
// sequence, different temporaries), there is no SEXP, no GC, no RNGScope, and none of the attribute machinery.
#ifndef ORACLE_RCPP_SHIM_H
#define ORACLE_RCPP_SHIM_H

#include <cmath>
#include <cstddef>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef enum { FALSE = 0, TRUE } Rboolean;   // R_ext/Boolean.h
#ifndef NA_LOGICAL
#define NA_LOGICAL (std::numeric_limits<int>::min())   // R_NaInt
#endif

namespace Rcpp {

struct index_out_of_bounds : public std::out_of_range {
  explicit index_out_of_bounds(const std::string& what) : std::out_of_range(what) {}
};
struct not_compatible : public std::runtime_error {
  explicit not_compatible(const std::string& what) : std::runtime_error(what) {}
};

struct Underscore {};
static const Underscore _ = Underscore();

static std::ostream& Rcerr = std::cerr;
static std::ostream& Rcout = std::cout;

// ---------------------------------------------------------------------------------------------- strings
class String {
  std::string s_;
 public:
  String() {}
  String(const char* s) : s_(s) {}
  String(const std::string& s) : s_(s) {}
  const std::string& get() const { return s_; }
  bool operator==(const String& o) const { return s_ == o.s_; }
  bool operator==(const char* o) const { return s_ == o; }
  bool operator!=(const String& o) const { return s_ != o.s_; }
  bool operator!=(const char* o) const { return s_ != o; }
};

class StringVector {
  std::shared_ptr<std::vector<String> > p_;
 public:
  StringVector() : p_(new std::vector<String>()) {}
  explicit StringVector(int n) : p_(new std::vector<String>(n)) {}
  int size() const { return (int)p_->size(); }
  int length() const { return size(); }
  String& operator[](std::ptrdiff_t i) { return p_->at((size_t)i); }
  const String& operator[](std::ptrdiff_t i) const { return p_->at((size_t)i); }
  void push_back(const String& s) { p_->push_back(s); }
};
typedef StringVector CharacterVector;

// ---------------------------------------------------------------------------------------------- logicals
class LogicalVector {
  std::shared_ptr<std::vector<int> > p_;
 public:
  explicit LogicalVector(int n) : p_(new std::vector<int>(n, 0)) {}
  int size() const { return (int)p_->size(); }
  int& operator[](std::ptrdiff_t i) { return (*p_)[(size_t)i]; }
  int operator[](std::ptrdiff_t i) const { return (*p_)[(size_t)i]; }
};

inline LogicalVector operator==(const LogicalVector& a, int b) {
  LogicalVector r(a.size());
  for (int i = 0; i < a.size(); i++) r[i] = (a[i] == NA_LOGICAL || b == NA_LOGICAL) ? NA_LOGICAL : (a[i] == b);
  return r;
}
// R's three-valued all()/any() (Rcpp::sugar::All / Any)
inline int all(const LogicalVector& v) {
  bool na = false;
  for (int i = 0; i < v.size(); i++) {
    if (v[i] == FALSE) return FALSE;
    if (v[i] == NA_LOGICAL) na = true;
  }
  return na ? NA_LOGICAL : TRUE;
}
inline int any(const LogicalVector& v) {
  bool na = false;
  for (int i = 0; i < v.size(); i++) {
    if (v[i] == TRUE) return TRUE;
    if (v[i] == NA_LOGICAL) na = true;
  }
  return na ? NA_LOGICAL : FALSE;
}
inline bool is_true(int x) { return x == TRUE; }
inline bool is_false(int x) { return x == FALSE; }

// ---------------------------------------------------------------------------------------------- numerics
class NumericVector {
  std::shared_ptr<std::vector<double> > p_;
 public:
  NumericVector() : p_(new std::vector<double>()) {}
  explicit NumericVector(int n) : p_(new std::vector<double>((size_t)n, 0.0)) {}
  NumericVector(const double* first, const double* last) : p_(new std::vector<double>(first, last)) {}
  static NumericVector scalar(double v) { NumericVector r(1); r[0] = v; return r; }
  int size() const { return (int)p_->size(); }
  int length() const { return size(); }
  double& operator[](std::ptrdiff_t i) { return (*p_)[(size_t)i]; }
  double operator[](std::ptrdiff_t i) const { return (*p_)[(size_t)i]; }
  const double* begin() const { return p_->data(); }
  const double* end() const { return p_->data() + p_->size(); }
  static bool is_na(double x) { return std::isnan(x); }   // R_isnancpp: NA_real_ and NaN
};

namespace shim {
inline void same_size(const NumericVector& a, const NumericVector& b) {
  if (a.size() != b.size()) throw not_compatible("rcpp_shim: sugar operands of different length");
}
template <typename F> inline NumericVector map1(const NumericVector& a, F f) {
  NumericVector r(a.size());
  for (int i = 0; i < a.size(); i++) r[i] = f(a[i]);
  return r;
}
template <typename F> inline NumericVector map2(const NumericVector& a, const NumericVector& b, F f) {
  same_size(a, b);
  NumericVector r(a.size());
  for (int i = 0; i < a.size(); i++) r[i] = f(a[i], b[i]);
  return r;
}
}  // namespace shim

#define RCPP_SHIM_BINOP(OP)                                                                             \
  inline NumericVector operator OP(const NumericVector& a, const NumericVector& b) {                    \
    return shim::map2(a, b, [](double x, double y) { return x OP y; });                                 \
  }                                                                                                     \
  inline NumericVector operator OP(const NumericVector& a, double b) {                                  \
    return shim::map1(a, [b](double x) { return x OP b; });                                             \
  }                                                                                                     \
  inline NumericVector operator OP(double a, const NumericVector& b) {                                  \
    return shim::map1(b, [a](double y) { return a OP y; });                                             \
  }
RCPP_SHIM_BINOP(+)
RCPP_SHIM_BINOP(-)
RCPP_SHIM_BINOP(*)
RCPP_SHIM_BINOP(/)
#undef RCPP_SHIM_BINOP

inline NumericVector operator-(const NumericVector& a) {
  return shim::map1(a, [](double x) { return -x; });
}
inline NumericVector exp(const NumericVector& a) { return shim::map1(a, [](double x) { return ::exp(x); }); }
inline NumericVector log(const NumericVector& a) { return shim::map1(a, [](double x) { return ::log(x); }); }
inline NumericVector sqrt(const NumericVector& a) { return shim::map1(a, [](double x) { return ::sqrt(x); }); }
inline NumericVector abs(const NumericVector& a) { return shim::map1(a, [](double x) { return ::fabs(x); }); }
inline NumericVector pow(const NumericVector& a, double e) {
  return shim::map1(a, [e](double x) { return ::pow(x, e); });
}
inline double sum(const NumericVector& a) {
  double r = 0.0;
  for (int i = 0; i < a.size(); i++) r += a[i];
  return r;
}
inline LogicalVector operator==(const NumericVector& a, const NumericVector& b) {
  shim::same_size(a, b);
  LogicalVector r(a.size());
  for (int i = 0; i < a.size(); i++)
    r[i] = (std::isnan(a[i]) || std::isnan(b[i])) ? NA_LOGICAL : (a[i] == b[i] ? TRUE : FALSE);
  return r;
}

class NumericMatrix {
  std::shared_ptr<std::vector<double> > p_;
  int nrow_, ncol_;
 public:
  NumericMatrix() : p_(new std::vector<double>()), nrow_(0), ncol_(0) {}
  NumericMatrix(int nrow, int ncol)
      : p_(new std::vector<double>((size_t)nrow * (size_t)ncol, 0.0)), nrow_(nrow), ncol_(ncol) {}
  NumericMatrix(int nrow, int ncol, const double* colmajor)
      : p_(new std::vector<double>(colmajor, colmajor + (size_t)nrow * (size_t)ncol)), nrow_(nrow), ncol_(ncol) {}
  int nrow() const { return nrow_; }
  int ncol() const { return ncol_; }
  int rows() const { return nrow_; }
  int cols() const { return ncol_; }
  double& operator()(int i, int j) { return (*p_)[(size_t)j * (size_t)nrow_ + (size_t)i]; }
  double operator()(int i, int j) const { return (*p_)[(size_t)j * (size_t)nrow_ + (size_t)i]; }
  // x(i, _): Rcpp hands out a strided row view; converted to a NumericVector it is a copy of the row.
  NumericVector operator()(int i, Underscore) const {
    NumericVector r(ncol_);
    for (int j = 0; j < ncol_; j++) r[j] = (*this)(i, j);
    return r;
  }
  const double* begin() const { return p_->data(); }
  static bool is_na(double x) { return std::isnan(x); }
};

// ---------------------------------------------------------------------------------------------- lists
// A list element: only numeric vectors ever travel through the lists of the two reference files.
class RObject {
  NumericVector v_;
 public:
  RObject() {}
  RObject(double x) : v_(NumericVector::scalar(x)) {}
  RObject(int x) : v_(NumericVector::scalar((double)x)) {}
  RObject(const NumericVector& v) : v_(v) {}
  const NumericVector& vec() const { return v_; }
  double as_double() const {
    if (v_.size() != 1) throw not_compatible("Expecting a single value: [extent=" + std::to_string(v_.size()) + "].");
    return v_[0];
  }
  operator double() const { return as_double(); }
  operator NumericVector() const { return v_; }
};

template <typename T> inline T as(const RObject& o);
template <> inline double as<double>(const RObject& o) { return o.as_double(); }
template <> inline NumericVector as<NumericVector>(const RObject& o) { return o.vec(); }

struct NamedValue {
  std::string name;
  RObject value;
};
class Named {
  std::string name_;
 public:
  explicit Named(const char* n) : name_(n) {}
  template <typename T> NamedValue operator=(const T& v) const {
    NamedValue nv;
    nv.name = name_;
    nv.value = RObject(v);
    return nv;
  }
};

class List {
  typedef std::vector<std::pair<std::string, RObject> > store_t;
  std::shared_ptr<store_t> p_;
  const RObject& find(const std::string& name) const {
    for (size_t i = 0; i < p_->size(); i++)
      if ((*p_)[i].first == name) return (*p_)[i].second;
    throw index_out_of_bounds("Index out of bounds: [index='" + name + "'].");
  }
 public:
  List() : p_(new store_t()) {}
  int size() const { return (int)p_->size(); }
  void push_back(const std::string& name, const RObject& v) { p_->push_back(std::make_pair(name, v)); }
  RObject operator[](const char* name) const { return find(name); }
  RObject operator[](const String& name) const { return find(name.get()); }
  static List create() { return List(); }
  template <typename... Rest> static List create(const NamedValue& first, const Rest&... rest) {
    List l;
    l.add(first, rest...);
    return l;
  }
 private:
  void add() {}
  template <typename... Rest> void add(const NamedValue& first, const Rest&... rest) {
    push_back(first.name, first.value);
    add(rest...);
  }
};

// ---------------------------------------------------------------------------------------------- R closures
namespace shim {
typedef std::function<RObject(const std::vector<RObject>&)> closure_t;
inline std::map<std::string, closure_t>& global_env() {
  static std::map<std::string, closure_t> env;
  return env;
}
}  // namespace shim

class Function {
  std::string name_;
 public:
  Function(const char* name) : name_(name) {
    // Rcpp::Function(name) looks the symbol up at construction and throws when it is not a function.
    if (shim::global_env().find(name_) == shim::global_env().end())
      throw std::runtime_error("rcpp_shim: no R function named '" + name_ + "' registered");
  }
  template <typename... Args> RObject operator()(const Args&... args) const {
    std::vector<RObject> a;
    collect(a, args...);
    return shim::global_env()[name_](a);
  }
 private:
  static void collect(std::vector<RObject>&) {}
  template <typename T, typename... Rest> static void collect(std::vector<RObject>& a, const T& t, const Rest&... rest) {
    a.push_back(RObject(t));
    collect(a, rest...);
  }
};

}  // namespace Rcpp

#endif  // ORACLE_RCPP_SHIM_H
