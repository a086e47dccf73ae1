"""The reference driver's command line (MC-GRA/main.py, README commands) end to end on a small synthetic graph."""
import os

import pytest

pytestmark = pytest.mark.gpu


def test_main_cli_readme_flags(tmp_path, monkeypatch):
    from mcgra_b200 import main as cli
    monkeypatch.chdir(tmp_path)                      # ./results/<log_name> is written relative to the cwd, as in the reference
    auc = cli.main(["--dataset", "synthetic:600:48:3", "--measure", "MSELoss", "--w1", "0.01", "--w6", "10", "--w7", "10",
                    "--w9", "10", "--w10", "1000", "--lr", "-2", "--useH_A", "--useY_A", "--useY", "--epochs", "20",
                    "--log_name", "cli_test.log"])
    assert 0.5 < float(auc) <= 1.0                   # priors H_A / Y_A / Y alone already beat chance on a planted partition
    log = open(os.path.join("results", "cli_test.log")).read()
    assert "In Whole Graph: AUC=" in log and "current density:" in log
