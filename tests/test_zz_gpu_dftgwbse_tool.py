"""The dftgwbse tool (votca_b200/tools/dftgwbse.py: the gwbse task of `xtp_tools -e dftgwbse`): the
reference's options file in, <job_name>.orb and <job_name>_summary.xml out, same numbers as the job facade."""
import os
import re

import numpy as np
import pytest

from tests.helpers import load_golden, methane_integrals

pytestmark = pytest.mark.gpu

OPTIONS = """<?xml version="1.0"?>
<options>
  <dftgwbse>
    <job_name>methane</job_name>
    <tasks>gwbse</tasks>
    <dftpackage><name>xtp</name></dftpackage>
    <gwbse>
      <tasks>gw,singlets</tasks>
      <ranges>full</ranges>
      <gw><mode>G0W0</mode><qp_grid_steps>601</qp_grid_steps><qp_grid_spacing>0.005</qp_grid_spacing>
          <mixing_order>0</mixing_order></gw>
      <bse><exctotal>3</exctotal><useTDA>true</useTDA><davidson><tolerance>lapack</tolerance></davidson></bse>
    </gwbse>
  </dftgwbse>
</options>
"""


def test_tool_reproduces_the_reference_gw_fixture_and_writes_both_outputs(tmp_path, monkeypatch):
    from oracle.orbfile import OrbFile
    from votca_b200.tools import dftgwbse
    g, m = load_golden(), methane_integrals()
    monkeypatch.chdir(tmp_path)
    (tmp_path / "opt.xml").write_text(OPTIONS)
    np.savez(tmp_path / "dft.npz", mos=g["gw/mo_eigenvectors"], mo_energies=g["inline/gw_mo_eigenvalues"], homo=4,
             vxc=g["gw/vxc"], ao3c=m["ao3c"], aux_overlap=m["S"], aux_coulomb=m["V"], dft_total_energy=-40.0)
    assert dftgwbse.main(["-o", "opt.xml", "--dft", "dft.npz"]) == 0
    assert os.path.exists("methane.orb") and os.path.exists("methane_summary.xml")
    qp = OrbFile("methane.orb").read("/QMdata/QPpert_energies").ravel()
    ref = np.diag(g["gw/ref"])  # test_gw.cc: G0W0 on this input
    assert np.abs(qp - ref).max() / np.abs(ref).max() < 1e-4
    xml = open("methane_summary.xml").read()
    assert "<singlets>" in xml or "singlet" in xml
    assert len(re.findall(r"<level ", xml)) >= 17


def test_tool_options_parsing_and_task_gate(tmp_path, capsys):
    from votca_b200.tools import dftgwbse
    p = tmp_path / "o.xml"
    p.write_text(OPTIONS.replace("<tasks>gwbse</tasks>", "<tasks>input,dft,parse</tasks>"))
    assert dftgwbse.read_tool_options(str(p)) == ("methane", ["input", "dft", "parse"])
    assert dftgwbse.main(["-o", str(p), "--dft", str(tmp_path / "missing.npz")]) == 0  # nothing to do, nothing read
    assert "only task" in capsys.readouterr().out
