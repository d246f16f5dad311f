"""not-gpu: the launch plan of a pass (host logic of rs_job_run).  A pass is cut into consecutive
segments whose team width shrinks as the pass fills in; whatever the sizes, the segments must tile [0, pass_end)
exactly once, in order -- a gap would leave visits unclaimed, an overlap would run them twice."""
import itertools

import pytest

from resynthesizer_b200 import api


def pass_ends(n):
    ends, e = [n, n], n            # lib/passes.h:78-91: n, n, then three quarters of the previous, integer arithmetic
    for _ in range(4):
        e = e * 3 // 4
        ends.append(e)
    return ends


@pytest.mark.parametrize("n", [1, 2, 63, 4096, 8192, 8193, 16384, 24577, 32768, 65536, 65537, 200000, 200001,
                               262144, 600000, 600001, 1 << 20, (1 << 22) + 5, (1 << 28) + 1])
def test_segments_tile_every_pass(built_lib, n, monkeypatch):
    for v in ("RS_TEAM_P0", "RS_TEAM_PN", "RS_SEG_P0"):
        monkeypatch.delenv(v, raising=False)
    for patch, ordered in itertools.product((1, 9, 15, 16, 30, 64), (False, True)):
        for p, end in enumerate(pass_ends(n)):
            if end == 0:
                continue
            plan = api.plan_pass(n, end, patch, p, ordered)
            assert 1 <= len(plan) <= 4
            ends = [e for e, _ in plan]
            assert ends == sorted(set(ends)) and ends[-1] == end and ends[0] > 0
            assert all(w in (1, 2, 4, 8) for _, w in plan)
            widths = [w for _, w in plan]
            assert widths == sorted(widths, reverse=True) and len(set(widths)) == len(widths)   # shrinking, merged
            if p > 0:
                assert len(plan) == 1              # later passes: one launch


def test_plans_of_the_baseline_configurations(built_lib, monkeypatch):
    for v in ("RS_TEAM_P0", "RS_TEAM_PN", "RS_SEG_P0"):
        monkeypatch.delenv(v, raising=False)
    # cfg2: 1 Mi points, 9 neighbours, shuffled: the short plan, then the throughput kernel
    assert api.plan_pass(1 << 20, 1 << 20, 9, 0) == [(8192, 8), (24576, 4), (65536, 2), (1 << 20, 1)]
    assert api.plan_pass(1 << 20, 1 << 20, 9, 1) == [(1 << 20, 1)]
    # 2048x2048 heal of a 1024x1024 hole, 30 neighbours: the long plan
    assert api.plan_pass(1 << 20, 1 << 20, 30, 0) == [(16384, 8), (65536, 4), (262144, 2), (1 << 20, 1)]
    # cfg1 (4096 points): latency kernel at 8 warps per visit in every pass
    assert api.plan_pass(4096, 4096, 30, 0) == [(4096, 8)] and api.plan_pass(4096, 3072, 30, 2) == [(3072, 8)]
    # cfg5 (65536 points): 8 warps in pass 0, 4 afterwards
    assert api.plan_pass(65536, 65536, 30, 0) == [(65536, 8)] and api.plan_pass(65536, 65536, 30, 1) == [(65536, 4)]
    # a sorted order (matchContextType 2-8) stays in latency mode for the whole of pass 0
    assert api.plan_pass(1 << 20, 1 << 20, 30, 0, ordered_visits=True) == [(1 << 20, 4)]


def test_plan_overrides(built_lib, monkeypatch):
    monkeypatch.delenv("RS_TEAM_PN", raising=False)
    monkeypatch.delenv("RS_SEG_P0", raising=False)
    monkeypatch.setenv("RS_TEAM_P0", "2")
    assert api.plan_pass(1 << 20, 1 << 20, 9, 0) == [(1 << 20, 2)]
    monkeypatch.delenv("RS_TEAM_P0")
    monkeypatch.setenv("RS_SEG_P0", "1000:8,5000:2,0:1")
    assert api.plan_pass(1 << 20, 1 << 20, 9, 0) == [(1000, 8), (5000, 2), (1 << 20, 1)]
    assert api.plan_pass(3000, 3000, 9, 0) == [(3000, 8)]     # small jobs never go below their base width
