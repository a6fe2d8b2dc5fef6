"""Known-answer checks of oracle/mini_r (the interpreter that runs the reference's R files): each snippet's expected
value is what R prints for it (R Language Definition / base help pages).  The semantics that could silently change
a number in R/*.R are covered: recycling, column-major matrices, drop of dimensions, 1-based and negative indexing,
`*` vs `%*%`, solve/chol conventions (upper factor), apply over rows, lists with partial `$` matching, lazy
arguments and `...`, substitute/eval/parse, ifelse laziness, while/break/next, integer sequences."""
import numpy as np
import pytest

from oracle.mini_r import interp as RI


@pytest.fixture(scope="module")
def I():
    return RI.Interp()


def val(I, src):
    return RI.to_py(I.run(src))


@pytest.mark.parametrize("src, expected", [
    ("1:3 + c(10, 20, 30)", [11, 22, 33]),
    ("c(1, 2, 3, 4) * c(10, 100)", [10, 200, 30, 400]),                        # recycling
    ("x <- c(5, 6, 7, 8); x[-1]", [6, 7, 8]),                                  # negative index drops
    ("x <- c(5, 6, 7, 8); x[c(TRUE, FALSE)]", [5, 7]),                         # logical index recycles
    ("x <- c(5, 6, 7); x[5] <- 1; x", [5, 6, 7, np.nan, 1]),                   # assignment extends with NA
    ("x <- numeric(); x[0]", []),
    ("seq(from = 0, to = 1, by = 0.25)", [0, 0.25, 0.5, 0.75, 1]),
    ("rep(c(1, 2), times = 2)", [1, 2, 1, 2]),
    ("rep(c(1, 2), each = 2)", [1, 1, 2, 2]),
    ("5 %/% 2 + 5 %% 2 + 2^3", [11]),
    ("-2^2", [-4]),                                                            # unary minus binds looser than ^
    ("sum(1:10) / length(1:10)", [5.5]),
    ("cumsum(c(1, 2, 3))", [1, 3, 6]),
    ("which(c(FALSE, TRUE, TRUE))", [2, 3]),
    ("which.max(c(3, 9, 9, 1))", [2]),                                         # first maximum
    ("ifelse(c(1, -1, 2) > 0, 1, 0)", [1, 0, 1]),
    ("ifelse(FALSE, numeric()[1], 7)", [7]),
    ("abs(sign(-3) - sign(2)) / 2", [1]),
    ("sqrt(c(4, 9)) * exp(0) + log(1)", [2, 3]),
    ("max(abs(c(-3, 2))) > 1e100 || any(is.nan(c(1, 2)))", [False]),
    ("plogis(q = 0, log.p = TRUE)", [np.log(0.5)]),
    ("plogis(-800, log.p = TRUE)", [-800.0]),                                  # no underflow to -Inf
])
def test_vector_semantics(I, src, expected):
    got = np.asarray(val(I, src), dtype=np.float64).reshape(-1)
    np.testing.assert_allclose(got, np.asarray(expected, dtype=np.float64), rtol=1e-15, equal_nan=True)


def test_matrices_are_column_major_and_drop(I):
    assert val(I, "matrix(1:6, nrow = 2)").tolist() == [[1, 3, 5], [2, 4, 6]]
    assert val(I, "matrix(1:6, nrow = 2, byrow = TRUE)").tolist() == [[1, 2, 3], [4, 5, 6]]
    assert val(I, "m <- matrix(1:6, nrow = 2); m[2, ]").tolist() == [2, 4, 6]              # a row drops to a vector
    assert val(I, "m <- matrix(1:6, nrow = 2); dim(m[, 2:3])").tolist() == [2, 2]
    assert val(I, "m <- matrix(1:6, nrow = 2); m[, 2] <- c(0, 0); m").tolist() == [[1, 0, 5], [2, 0, 6]]
    assert val(I, "m <- matrix(0, 2, 2); m[2, 1] <- 5; t(m)").tolist() == [[0, 5], [0, 0]]
    assert val(I, "dim(matrix())").tolist() == [1, 1] and np.isnan(val(I, "matrix()")[0, 0])
    assert val(I, "nrow(diag(3)) + ncol(diag(3))").tolist() == [6]
    assert val(I, "diag(matrix(1:4, 2))").tolist() == [1, 4]
    assert val(I, "as.numeric(matrix(1:4, 2))").tolist() == [1, 2, 3, 4]


def test_elementwise_vs_matrix_product_and_row_scaling(I):
    # (1/Z) * Sigma12 scales ROWS (the vector recycles down the columns): used on every line of the reference
    assert val(I, "c(1, 10) * matrix(1, 2, 3)").tolist() == [[1, 1, 1], [10, 10, 10]]
    assert val(I, "matrix(1:4, 2) %*% matrix(1:4, 2)").tolist() == [[7, 15], [10, 22]]
    assert val(I, "t(1:3) %*% 1:3").tolist() == [[14]]                                     # vector %*% vector
    assert val(I, "dim(1:3 %*% t(1:3))").tolist() == [3, 3]
    assert val(I, "apply(X = matrix(1:6, 2), MARGIN = 1, FUN = sum)").tolist() == [9, 12]
    assert val(I, "rowSums(matrix(1:6, 2)) - colSums(t(matrix(1:6, 2)))").tolist() == [0, 0]


def test_linear_algebra_conventions(I):
    I.run("A <- matrix(c(4, 2, 2, 3), 2)")
    R = val(I, "chol(A)")
    assert R[1, 0] == 0.0                                                                  # chol() is UPPER: t(R) %*% R = A
    np.testing.assert_allclose(R.T @ R, [[4, 2], [2, 3]], rtol=1e-15)
    np.testing.assert_allclose(val(I, "solve(A)"), np.linalg.inv([[4, 2], [2, 3]]), rtol=1e-14)
    np.testing.assert_allclose(val(I, "solve(a = A, b = c(1, 2))").reshape(-1), np.linalg.solve([[4, 2], [2, 3]], [1, 2]), rtol=1e-14)
    np.testing.assert_allclose(val(I, "det(A)"), [8.0], rtol=1e-14)
    np.testing.assert_allclose(val(I, "2 * sum(log(diag(chol(A))))"), [np.log(8.0)], rtol=1e-14)
    with pytest.raises(RI.RError):
        I.run("chol(matrix(c(1, 2, 2, 1), 2))")                                            # not positive definite -> error (try() relies on it)


def test_lists_functions_and_lazy_arguments(I):
    assert val(I, 'l <- list("sigma" = 2, "tau" = 3); l$tau * l[["sigma"]]').tolist() == [6]
    assert val(I, 'l <- list(learn_rate = 4); l$learn').tolist() == [4]                    # `$` partial matching
    assert val(I, 'l <- list(a = 1); is.null(l$b)').tolist() == [True]
    assert val(I, 'names(list(a = 1, b = 2))').tolist() == ["a", "b"]
    assert val(I, 'unlist(list(a = 1, b = 2)) + 1') == {"a": 2.0, "b": 3.0}
    assert val(I, 'f <- function(x, y = x * 2) { x <- 10; y }; f(1)').tolist() == [20]     # default forced lazily
    assert val(I, 'f <- function(x, unused) x; f(3)').tolist() == [3]                      # missing arg never forced
    assert val(I, 'f <- function(...) { a <- list(...); a$m }; f(m = 7)').tolist() == [7]
    assert val(I, 'g <- function(z, ...) h(z, ...); h <- function(z, k = 1) z * k; g(2, k = 5)').tolist() == [10]
    assert val(I, 'f <- function() { for(i in 1:5) { if(i == 2) next; if(i == 4) break }; i }; f()').tolist() == [4]
    assert val(I, 'k <- 0; while(TRUE) { k <- k + 1; if(k >= 3) break }; k').tolist() == [3]
    assert val(I, 'is.function(sum) && !is.function(NA) && is.list(list()) && !is.list(NA)').tolist() == [True]
    assert val(I, 'do.call(what = "sum", args = list(1, 2, 3))').tolist() == [6]
    assert val(I, 'sapply(1:3, function(i) i^2)').tolist() == [1, 4, 9]


def test_substitute_eval_parse_as_the_reference_uses_them(I):
    # R/covariance_function_derivatives.R:249: eval(parse(text = eval(substitute(paste("cov_par$l", b, sep = ""), list(b = i)))))
    I.run('cov_par <- list(sigma = 1, l1 = 0.5, l2 = 0.25)')
    assert val(I, 'i <- 2; eval(parse(text = eval(substitute(expr = paste("cov_par$l", b, sep = ""), env = list("b" = i)))))').tolist() == [0.25]
    # R/vi_functions.R:441: eval(substitute(f(x1 = a), list(a = value)))
    assert val(I, 'f <- function(x1) x1 + 1; eval(substitute(expr = f(x1 = a), env = list("a" = 41)))').tolist() == [42]
    assert val(I, 'paste("l", 1:3, sep = "")').tolist() == ["l1", "l2", "l3"]
    assert val(I, '"l2" %in% paste("l", 1:3, sep = "")').tolist() == [True]


def test_namespaced_helpers_of_the_reference(I):
    assert val(I, "m <- Matrix::Matrix(data = 0, nrow = 2, ncol = 3); m[, 2] <- c(1, 2); m").tolist() == [[0, 1, 0], [0, 2, 0]]
    a = val(I, "abind::abind(abind::abind(matrix(1:4, 2), matrix(5:8, 2), along = 3), matrix(9:12, 2), along = 3)")
    assert a.shape == (2, 2, 3) and a[:, :, 2].tolist() == [[9, 11], [10, 12]]
