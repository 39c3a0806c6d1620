"""ORACLE (test infrastructure, not product code): sigma schedule, EDM preconditioning and
sampler-step algebra, restated in plain Python floats (fp64).

PARITY UNPINNED: /root/reference ships no code, tests or golden vectors
(/root/reference/README.md:11-16 "Under construction"); these formulas follow the pinned spec in
SURVEY.md §8(a) rows a-1, a-2, a-6 (Karras et al. 2022 "EDM" schedule/preconditioning and the
ADPM2 ancestral step the StyleTTS lineage uses).  They are pinned by analytic known-answer tests in
tests/test_oracle_kat.py instead.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.
"""
from __future__ import annotations

import math
from typing import List, Tuple


def karras_sigmas(n: int, sigma_min: float, sigma_max: float, rho: float) -> List[float]:
    """a-1: sigma_i = (smax^(1/rho) + i/(n-1) (smin^(1/rho) - smax^(1/rho)))^rho, i = 0..n-1.
    n == 1 -> [sigma_max]."""
    if n == 1:
        return [float(sigma_max)]
    a, b = sigma_max ** (1.0 / rho), sigma_min ** (1.0 / rho)
    return [(a + i / (n - 1) * (b - a)) ** rho for i in range(n)]


def student_sigmas(steps: int, cfg) -> List[float]:
    """Distilled student: `steps` Karras points followed by a terminal 0 (steps transitions)."""
    return karras_sigmas(steps, cfg.sigma_min, cfg.sigma_max, cfg.rho) + [0.0]


def teacher_sigmas(steps: int, cfg) -> List[float]:
    """Teacher: steps+1 Karras points, `steps` ADPM2 transitions, ends at sigma_min."""
    return karras_sigmas(steps + 1, cfg.sigma_min, cfg.sigma_max, cfg.rho)


def edm_precond(sigma: float, sigma_data: float) -> Tuple[float, float, float, float]:
    """a-2: (c_skip, c_out, c_in, c_noise)."""
    s2, d2 = sigma * sigma, sigma_data * sigma_data
    c_skip = d2 / (s2 + d2)
    c_out = sigma * sigma_data / math.sqrt(s2 + d2)
    c_in = 1.0 / math.sqrt(s2 + d2)
    c_noise = math.log(sigma) / 4.0
    return c_skip, c_out, c_in, c_noise


def adpm2_sigmas(sigma: float, sigma_next: float) -> Tuple[float, float, float]:
    """a-6: (sigma_up, sigma_down, sigma_mid) of one ADPM2 step, rho = 1 midpoint."""
    sigma_up = math.sqrt(sigma_next ** 2 * (sigma ** 2 - sigma_next ** 2) / sigma ** 2)
    sigma_down = math.sqrt(max(sigma_next ** 2 - sigma_up ** 2, 0.0))
    sigma_mid = 0.5 * (sigma + sigma_down)
    return sigma_up, sigma_down, sigma_mid


def time_features(c_noise: float, d_time: int) -> List[float]:
    """Sinusoidal features of c_noise: [sin(c f_i) | cos(c f_i)], f_i = 100^(i/(half-1))."""
    half = d_time // 2
    fr = [math.exp(math.log(100.0) * i / max(half - 1, 1)) for i in range(half)]
    return [math.sin(c_noise * f) for f in fr] + [math.cos(c_noise * f) for f in fr]
