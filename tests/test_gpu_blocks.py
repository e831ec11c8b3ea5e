"""GPU: the device building blocks of the assembly (prefix scan, stable radix sort) against numpy."""
import ctypes as C

import numpy as np
import pytest

from magnetite_b200 import _lib

pytestmark = pytest.mark.gpu


def gpu_scan(ctx, a):
    out = np.empty(a.size + 1, np.uint32)
    _lib.check(_lib.load().mag_debug_exclusive_scan(ctx.handle, _lib.ptr(a), _lib.ptr(out), a.size), "scan")
    return out


def gpu_sort(ctx, keys, pay, bits):
    k, p = keys.copy(), pay.copy()
    _lib.check(_lib.load().mag_debug_sort_pairs(ctx.handle, _lib.ptr(k), _lib.ptr(p), k.size, bits), "sort")
    return k, p


@pytest.mark.parametrize("n", [0, 1, 31, 2047, 2048, 2049, 100_000, 2048 * 2048 + 17])
def test_exclusive_scan(ctx, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 5, n).astype(np.uint32)
    out = gpu_scan(ctx, a)
    ref = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)]).astype(np.uint32)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("n,bits", [(2, 8), (4095, 16), (4096, 24), (4097, 40), (300_000, 46), (1_000_003, 50)])
def test_radix_sort_is_correct_and_stable(ctx, n, bits):
    rng = np.random.default_rng(n + bits)
    # few distinct keys -> long runs of equal keys -> stability is exercised
    distinct = rng.integers(0, 1 << bits, max(2, n // 7), dtype=np.uint64)
    keys = distinct[rng.integers(0, distinct.size, n)]
    pay = np.arange(n, dtype=np.uint32)
    k, p = gpu_sort(ctx, keys, pay, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(p, pay[order].astype(np.uint32))


def test_radix_sort_ignores_high_bits_and_handles_sorted_input(ctx):
    n = 50_000
    keys = (np.arange(n, dtype=np.uint64) // 3) | (np.uint64(1) << np.uint64(60))
    pay = np.arange(n, dtype=np.uint32)
    k, p = gpu_sort(ctx, keys, pay, 20)
    assert np.array_equal(k, keys) and np.array_equal(p, pay)
    rev = keys[::-1].copy()
    k, p = gpu_sort(ctx, rev, pay, 20)
    order = np.argsort(rev & np.uint64((1 << 20) - 1), kind="stable")
    assert np.array_equal(k, rev[order]) and np.array_equal(p, pay[order])
