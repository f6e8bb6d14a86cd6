"""Result fusion mirror (mevi_b200/ensemble.py) against report text produced by the UNMODIFIED reference
scripts (tests/golden/make_ensemble_golden.py ran MEVI/ensemble_marco.py and ensemble_nqdpr.py)."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def _run(kind, tmp_path):
    import make_ensemble_golden as g

    from mevi_b200 import ensemble

    work = str(tmp_path / "in")
    g.ensemble_inputs(work)
    ofile = str(tmp_path / "report.txt")
    args = g.marco_args(work, ofile) if kind == "marco" else g.nq_args(work, ofile)
    args.device = "host"  # the reference's python dictionaries; the device path is tests/test_gpu_ensemble.py
    if kind == "marco":
        ensemble.combine_main_marco(args)
    else:
        ensemble.combine_main_nqdpr(args)
    return open(ofile).read(), work


@pytest.mark.parametrize("kind", ["marco", "nqdpr"])
def test_report_text_equals_reference(kind, tmp_path, capsys):
    text, work = _run(kind, tmp_path)
    golden = open(os.path.join(HERE, "golden", "ensemble", f"ensemble_{kind}_report.txt")).read()
    assert text == golden
    # the parse / rank caches the reference writes next to its inputs exist and are reused on a second run
    for f in ("ance.pkl", "fine.pkl", "coarse.pkl", "coarse_cr4gt.pkl", "fine_cr.pkl"):
        assert os.path.exists(os.path.join(work, f)), f
    import make_ensemble_golden as g

    from mevi_b200 import ensemble

    ofile2 = str(tmp_path / "report2.txt")
    args = g.marco_args(work, ofile2) if kind == "marco" else g.nq_args(work, ofile2)
    args.device = "host"
    if kind == "marco":
        ensemble.combine_main_marco(args)
    else:
        ensemble.combine_main_nqdpr(args)
    assert open(ofile2).read() == golden


def test_fuse_arithmetic():
    from mevi_b200.ensemble import cluster_rankings, fuse, ranking_of

    mapping = {1: (0, 0), 2: (1, 1), 3: (2, 2)}
    ranks, num = cluster_rankings({"q": [1, 2, 3, -1]}, {"q": [[1, 1], [0, 0]]}, mapping)
    assert ranks["q"] == [1, 0, 2, 2] and num == 2
    f = fuse([1, 2, 3, -1], [1.0, 1.0, 1.0, 5.0], ranks["q"], 0.6, 0.03, 0.02, num)
    assert f[2] == 1.0 + 0.6 / 1.0 and f[1] == 1.0 + 0.6 / 1.03
    assert f[3] == (1.0 + 0.6 / 1.06) * (1 - 0.02 * 0.6)
    assert ranking_of(f)[0] == -1 and ranking_of(f)[1] == 2


def test_mapping_to_codes_marks_holes():
    import numpy as np

    from mevi_b200.ensemble import mapping_to_codes

    codes = mapping_to_codes({0: (1, 2, 3), 3: (0, 0, 1), 1: (2, 2, 2)})
    assert codes.shape == (4, 3) and codes.dtype == np.int32
    assert codes[0].tolist() == [1, 2, 3] and codes[1].tolist() == [2, 2, 2] and codes[3].tolist() == [0, 0, 1]
    assert (codes[2] == np.iinfo(np.int32).min).all()  # document 2 is not a key: the rank kernel reports a KeyError for it


def test_device_path_is_the_default_and_fails_loudly_without_a_gpu(tmp_path):
    """`--device cuda` (the CLI default) never falls back to the python dictionaries: without a CUDA device the drivers
    raise; the host arithmetic runs only when asked for by name."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import make_ensemble_golden as g

    import mevi_b200
    from mevi_b200 import ensemble

    work = str(tmp_path / "in")
    g.ensemble_inputs(work)
    args = g.marco_args(work, str(tmp_path / "r.txt"))  # no `device` attribute: the default applies
    with pytest.raises(mevi_b200.MeviError):
        ensemble.combine_main_marco(args)
    args = g.nq_args(work, str(tmp_path / "r2.txt"))
    args.device = "cuda"
    with pytest.raises(mevi_b200.MeviError):
        ensemble.combine_main_nqdpr(args)
    args = g.nq_args(work, str(tmp_path / "r3.txt"))  # (the drivers parse the flag strings in place, like the reference)
    args.device = "tpu"
    with pytest.raises(ValueError):
        ensemble.combine_main_nqdpr(args)
