"""The gpu_benchmark tool (votca_b200/tools/gpu_benchmark.py), the counterpart of the reference's
`xtp_tools -e gpu_benchmark` (xtp/src/libxtp/tools/gpu_benchmark.cc): statistics / output format on the CPU, one small
run on the GPU."""
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from votca_b200.tools import gpu_benchmark as gb

NAMES = ["Filling_ThreeCenter", "Multiplication_of_tensor_with_matrix", "RPA_evaluation", "SingletOperator_TDA",
         "TripletOperator_TDA", "SingletOperator_BTDA_B", "HxOperator"]


def test_statistics_and_xml_layout():
    mean, std = gb.calc_statistics([1.0, 2.0, 3.0, 4.0])  # population std, gpu_benchmark.cc:45-50
    assert mean == 2.5 and abs(std - 1.25 ** 0.5) < 1e-15
    calls = []
    part = gb.run_part(lambda: calls.append(1), "RPA_evaluation", 3, lambda: None)
    assert part[0] == "RPA_evaluation" and len(part[3]) == 3 and len(calls) == 3
    root = ET.fromstring(gb.to_xml({"Repetitions": 3, "GPUs": 1}, [part]))
    assert root.tag == "GPU_Benchmark" and root.find("Repetitions").text == "3"
    node = root.find("RPA_evaluation")
    assert float(node.find("avg").text) >= 0.0 and len(node.find("runs").findall("timing")) == 3
    # the operator template arguments of bse_operator.h:79-87
    assert dict(gb.OPERATORS) == {"SingletOperator_TDA": (1, 2, 1, 0), "TripletOperator_TDA": (1, 0, 1, 0),
                                  "SingletOperator_BTDA_B": (0, 2, 0, 1), "HxOperator": (0, 1, 0, 0)}


class _RecordingContext:
    """Stands in for votca_b200.api.Context on the CPU: records the call sequence, returns arrays of the right shape.
    It checks the tool's own plumbing (argument order, shapes, part sequence), nothing numerical."""

    def __init__(self, device=0):
        self.calls = []

    def __getattr__(self, name):
        def method(*args):
            self.calls.append(name)
            if name == "mmn_alloc":
                self.naux = args[0]
            if name == "pseudo_invsqrt":
                return np.eye(args[0].shape[0]), 0
            if name == "rpa_epsilon":
                return np.eye(self.naux)
            if name == "bse_configure":
                homo, rpamin, vmin, cmax, eps_inv, hqp = args
                assert eps_inv.shape == (self.naux,) and hqp.shape == (cmax - vmin + 1, cmax - vmin + 1)
                self.bse_size = (homo - vmin + 1) * (cmax - homo)
            if name == "bse_matmul":
                assert args[1].shape[0] == self.bse_size
                return np.zeros_like(args[1])
            return None
        return method


def test_tool_plumbing_on_the_cpu(tmp_path, monkeypatch):
    import votca_b200.api as api
    made = []

    def factory(device=0):
        made.append(_RecordingContext(device))
        return made[-1]
    monkeypatch.setattr(api, "Context", factory)
    out = tmp_path / "gb.xml"
    parts = gb.main(["--workload", "tiny", "--repetitions", "2", "--outputfile", str(out), "--spacesize", "3"])
    assert [p[0] for p in parts] == NAMES and len(parts[0][3]) == 2
    calls = made[0].calls
    assert calls.count("bse_matmul") == 8 and calls.count("rpa_epsilon") == 4 and calls.count("mmn_fill_block") == 3 * 2
    assert calls[-1] == "close" and ET.parse(out).getroot().find("HxOperator") is not None


@pytest.mark.gpu
def test_small_run_writes_all_parts(tmp_path):
    out = tmp_path / "gpu_benchmark.xml"
    parts = gb.main(["--workload", "tiny", "--repetitions", "2", "--outputfile", str(out), "--spacesize", "7"])
    assert [p[0] for p in parts] == NAMES
    root = ET.parse(out).getroot()
    assert [c.tag for c in root if c.find("avg") is not None] == NAMES
    assert all(float(root.find(n).find("avg").text) > 0.0 for n in NAMES)


@pytest.mark.gpu
def test_small_run_from_basis_sets(tmp_path):
    """--system: Filling_ThreeCenter includes the integrals (produced on the device), as the reference's Fill does."""
    out = tmp_path / "gpu_benchmark.xml"
    parts = gb.main(["--system", "methane-svp", "--repetitions", "1", "--outputfile", str(out), "--spacesize", "5"])
    assert [p[0] for p in parts] == NAMES
    root = ET.parse(out).getroot()
    assert root.find("Basisset").text == "def2-svp" and root.find("AuxBasissetsize").text == "104"
