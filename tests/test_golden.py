"""Golden vectors of the reference's own regression suite (tests/golden/golden_cases.*: 140 command lines of
the reference's tests/runtest with decoded inputs and expected outputs).

  * not-gpu: the reference build reproduces them (pins the checker), and so does the host-emulation build of
    this library (pins the host logic: plan, weight lists, colour tables, API mapping);
  * gpu: the CUDA library reproduces them bit-exactly through the C ABI.
"""
import pytest

import harness
from imageworsener_b200.jobs import Job

CASES = harness.load_golden()
IDS = [c["name"] for c in CASES]


def _run(lib, case):
    job = Job.parse(case["args"], outfmt=case["outfmt"], disable_all_opt=True)
    res = harness.run_job(lib, case["input"], job)
    assert res["ok"], res["err"]
    return harness.compare_with_golden(res, case)


def test_fixture_count():
    assert len(CASES) >= 130


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_reference_reproduces_golden(ref, case):
    assert _run(ref, case) == 0


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_host_logic_reproduces_golden(emu, case):
    assert _run(emu, case) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_cuda_reproduces_golden(prod, case):
    assert _run(prod, case) == 0
