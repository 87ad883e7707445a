"""CPU checks of the persistent kernels' tile sequences through the C ABI (l32_debug_tile_order): every tile exactly once,
and in the one-kernel tensor-parallel feed-forward every down tile comes well after the gate/up tiles whose act it reads."""
import ctypes

import pytest

from llama32_b200 import _lib


def _seq(kind, cfg, n):
    L = _lib.lib()
    arr = (ctypes.c_int * len(cfg))(*cfg)
    out = (ctypes.c_int * 3)()
    res = []
    for t in range(n):
        assert L.l32_debug_tile_order(kind, t, arr, out) == 0
        res.append((out[0], out[1], out[2]))
    return res


@pytest.mark.parametrize("tiles_m,tiles_n,group,rot", [(32, 112, 16, 0), (32, 16, 8, 5), (7, 3, 4, 2), (1, 9, 8, 0), (33, 5, 8, 32)])
def test_plain_order_is_a_permutation(tiles_m, tiles_n, group, rot):
    s = _seq(0, [tiles_m, tiles_n, group, rot, 0, 0, 0], tiles_m * tiles_n)
    assert sorted((m, n) for _, m, n in s) == [(m, n) for m in range(tiles_m) for n in range(tiles_n)]


@pytest.mark.parametrize("world,tpc,tiles_n,rank", [(8, 4, 16, 0), (8, 4, 16, 7), (2, 16, 16, 1), (4, 1, 3, 2)])
def test_reduce_scatter_order_round_robins_over_the_owners(world, tpc, tiles_n, rank):
    tiles_m = world * tpc
    s = _seq(0, [tiles_m, tiles_n, 8, 0, world, tpc, rank], tiles_m * tiles_n)
    assert sorted((m, n) for _, m, n in s) == [(m, n) for m in range(tiles_m) for n in range(tiles_n)]
    first_m = []                                      # m-tiles in first-visit order
    for _, m, _ in s:
        if m not in first_m:
            first_m.append(m)
    owners = [m // tpc for m in first_m]
    for i in range(0, len(owners) - world + 1, world):
        rnd = owners[i:i + world]
        assert sorted(rnd) == list(range(world)), "every round of m-tiles touches every owner once"
        assert rnd[-1] == rank, "the own rank (no NVLink traffic) comes last in each round"


@pytest.mark.parametrize("tiles_m,n_gu,n_dn,group,rot,prefix,clusters", [
    (32, 16, 16, 4, 8, 64, 74), (128, 16, 16, 8, 0, 74, 74), (32, 112, 16, 32, 0, 74, 74), (32, 28, 32, 4, 12, 74, 74),
    (6, 3, 2, 2, 1, 4, 4), (5, 7, 3, 1, 0, 7, 3)])
def test_ffn_order_visits_everything_once_and_respects_dependencies(tiles_m, n_gu, n_dn, group, rot, prefix, clusters):
    total = tiles_m * (n_gu + n_dn)
    s = _seq(1, [tiles_m, n_gu, n_dn, group, rot, prefix], total)
    gu = sorted((m, n) for p, m, n in s if p == 0)
    dn = sorted((m, n) for p, m, n in s if p == 1)
    assert gu == [(m, n) for m in range(tiles_m) for n in range(n_gu)]
    assert dn == [(m, n) for m in range(tiles_m) for n in range(n_dn)]
    last_gu = {}
    for t, (p, m, _) in enumerate(s):
        if p == 0:
            last_gu[m] = t
    dist = [t - last_gu[m] for t, (p, m, _) in enumerate(s) if p == 1]
    assert min(dist) >= 1, "a down tile may only wait for tiles that come earlier in the sequence (no deadlock)"
    if tiles_m >= 4 * group and group * n_gu >= clusters:
        assert min(dist) > clusters, "act tiles finish more than one wave before the down tiles that read them"
