"""Seeded synthetic cases shaped like BASELINE.json's configs."""
import numpy as np


def ard_par(sigma, ls, tau):
    cp = {"sigma": float(sigma)}
    for i, l in enumerate(ls):
        cp["l%d" % (i + 1)] = float(l)
    cp["tau"] = float(tau)
    return cp


def lvec(cp):
    return [cp[k] for k in cp if k.startswith("l")]


def config1(n=500):
    """1-D sqexp, exec/simulated_examples.R shape: x sorted U(0,10), theta = (2, 1, 1), knots (1,3,5,7,9)."""
    rng = np.random.default_rng(1308)
    x = np.sort(rng.uniform(0, 10, n)).reshape(-1, 1)
    cp = {"sigma": 2.0, "l": 1.0, "tau": 1.0}
    f = 2.0 * np.sin(x[:, 0]) + 0.5 * np.cos(2.3 * x[:, 0])
    y = f + rng.normal(0, cp["tau"], n)
    xu = np.array([1.0, 3.0, 5.0, 7.0, 9.0]).reshape(-1, 1)
    return dict(x=x, y=y, mu=np.zeros(n), xu=xu, cov_par=cp, cov_fun="sqexp", delta=1e-6)


def config2(n=1503, d=5, m=64):
    """airfoil-shaped: standardised N(0, I_5) inputs, knots = random rows + N(0, 0.01^2) jitter, ARD."""
    rng = np.random.default_rng(1309)
    x = rng.normal(size=(n, d))
    y = np.sin(x[:, 0]) + 0.5 * x[:, 1] - 0.3 * x[:, 2] * x[:, 3] + 0.3 * rng.normal(size=n)
    xu = x[rng.choice(n, m, replace=False)] + 0.01 * rng.normal(size=(m, d))
    s = float(np.sqrt(np.var(y) / 2))
    cp = ard_par(s, [1.0] * d, s)
    return dict(x=x, y=y, mu=np.zeros(n), xu=xu, cov_par=cp, cov_fun="ard", delta=1e-4)


def config3(n=4784, d=4, m=256, extra_knot_is_data_row=True):
    """ccpp-shaped OAT step: m knots + one candidate knot that IS a data row (quirk Q4)."""
    rng = np.random.default_rng(1310)
    x = rng.normal(size=(n, d))
    y = 0.8 * x[:, 0] - 0.4 * np.tanh(x[:, 1]) + 0.2 * x[:, 2] * x[:, 3] + 0.25 * rng.normal(size=n)
    xu = x[rng.choice(n, m, replace=False)] + 0.05 * rng.normal(size=(m, d))
    if extra_knot_is_data_row:
        xu = np.vstack([xu, x[rng.integers(n)]])
    cp = ard_par(1.0, [1.2, 0.9, 1.5, 1.1], 0.3)
    return dict(x=x, y=y, mu=np.zeros(n), xu=xu, cov_par=cp, cov_fun="ard", delta=1e-3)


def config4(n=2000, d=8, m=64):
    """Bernoulli classification shape (reduced n, m for the CPU oracle)."""
    rng = np.random.default_rng(1311)
    x = rng.normal(size=(n, d))
    f = 1.5 * np.sin(x[:, 0]) + x[:, 1] - 0.5 * x[:, 2]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(np.float64)
    xu = rng.normal(size=(m, d))
    cp = ard_par(2.0, [1.5] * d, 0.1)
    return dict(x=x, y=y, mu=np.zeros(n), xu=xu, cov_par=cp, cov_fun="ard", delta=1e-3)


def config5(n=4096, d=8, m=1024, seed=1312):
    """Headline recipe at reduced n: X, U ~ N(0, I_8), sigma 1, l_c = 0.8 + 0.05 c, tau 0.5, delta 1e-6."""
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, d))
    xu = rng.normal(size=(m, d))
    cp = ard_par(1.0, [0.8 + 0.05 * (c + 1) for c in range(d)], 0.5)
    y = np.sin(x[:, 0]) + 0.5 * x[:, 1] + cp["tau"] * rng.normal(size=n)
    return dict(x=x, y=y, mu=np.zeros(n), xu=xu, cov_par=cp, cov_fun="ard", delta=1e-6)
