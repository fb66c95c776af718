"""The plan fixtures against the only golden vector the reference holds for this path:
the Q6 excerpt in its README (README.md:40-52), extracted by tools/make_golden.py."""
import os

from util import ROOT, plan_text


def test_q06_matches_readme_golden_lines():
    golden = open(os.path.join(ROOT, "tests", "golden", "readme_q06_lines.txt")).read().splitlines()
    plan = plan_text("q06.vdl").splitlines()
    assert "..." in golden
    head = golden[: golden.index("...")]
    tail = golden[golden.index("...") + 1:]
    assert len(head) == 9 and len(tail) == 3
    assert plan[: len(head)] == head
    assert plan[-len(tail):] == tail
    # the README's last statement id pins the statement count at 42
    assert int(tail[-1].split(",")[0]) == len(plan) == 42


def test_plans_are_numbered_and_backward_referencing():
    import re
    for name in ("q06.vdl", "q01.vdl"):
        for i, line in enumerate(plan_text(name).splitlines(), 1):
            assert line.startswith(f"{i},"), (name, line)
            for ref in re.findall(r"Id (\d+)", line):
                assert 0 < int(ref) < i
