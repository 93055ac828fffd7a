"""Deterministic input builders shared by the hashed-golden tests (mirror oracle/gen_golden.py)."""
import numpy as np
import torch


def gen_normal(shape, sigma, seed):
    return torch.from_numpy((np.random.RandomState(seed).standard_normal(shape) * sigma).astype(np.float32))


def gen_softmax(shape, seed):
    s = torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32)) * 3
    n = shape[-1]
    mask = torch.triu(torch.ones(shape[-2], n, dtype=torch.bool), diagonal=1)
    s = s.masked_fill(mask, torch.finfo(torch.float32).min)
    return torch.softmax(s, dim=-1)


def hashed_input(case):
    shape = tuple(case["shape"])
    if case["sigma"] is None:
        return gen_softmax(shape, case["seed"])
    return gen_normal(shape, case["sigma"], case["seed"])
