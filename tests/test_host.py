"""CPU: host-side mirror of the reference interface, generators, and the C-ABI boundary (no compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, gold_npy


def test_library_exports_every_declared_symbol():
    from keff_cfd_b200 import _lib
    header = open(os.path.join(ROOT, "include", "keffb200.h")).read()
    declared = set(re.findall(r"\b(keffb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in _lib.lib().keffb200_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from keff_cfd_b200 import _lib
    p = _lib.Params()
    _lib.lib().keffb200_default_params(ctypes.byref(p))
    assert (p.ks, p.kf, p.tl, p.tr, p.convergence, p.max_iter) == (10, 1, 0, 1, 1e-5, 500000)
    assert _lib.lib().keffb200_init(0) == -1                       # KEFFB200_ENOGPU
    assert b"no CPU path" in _lib.lib().keffb200_last_error()
    r = _lib.Result()
    m = np.zeros((4, 4), dtype=np.uint8)
    rc = _lib.lib().keffb200_solve2d(m.ctypes.data, 4, 4, ctypes.byref(p), None, None, None, None, ctypes.byref(r))
    assert rc == -6                                                # KEFFB200_ESTATE: never computes on the CPU
    with pytest.raises(_lib.KeffB200Error):
        _lib.init(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "keff_cfd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower(), f


def test_read_input_file_reference_format(tmp_path):
    from keff_cfd_b200 import read_input_file
    txt = ("Input File:\nks: 10\nkf: 1\nWidth: 128\nHeight: 128\nDepth: 128\nMeshAmpX: 2\nMeshAmpY: 3\nMeshAmpZ: 1\n"
           "InputName: Schwarz163.csv\nTR: 1\nTL: 0\nOutputName: outputSample.csv\nprintTMap: 1\n"
           "TMapName: Schwarz163_T.csv\nprintQMap: 0\nQMapName: q.csv\nConvergence: 0.00001\nMaxIter: 500000\n"
           "Verbose: 1\nNumCores: 0\nMeshFlag: 0\nMaxMesh: 10\nRunBatch: 0\nNumImages: 500\nunknown: 3\n")
    p = tmp_path / "input.txt"
    p.write_text(txt)
    o = read_input_file(str(p))
    assert (o.TCsolid, o.TCfluid, o.Width, o.Height, o.Depth) == (10.0, 1.0, 128, 128, 128)
    assert (o.MeshIncreaseX, o.MeshIncreaseY, o.inputFilename, o.outputFilename) == (2, 3, "Schwarz163.csv", "outputSample.csv")
    assert (o.ConvergeCriteria, o.MAX_ITER, o.printTmap, o.TMapName, o.NumImg) == (1e-5, 500000, 1, "Schwarz163_T.csv", 500)
    assert o.numCores == 1  # "Entered number of cores not valid. Default = 1." (keff2D.h NumCores branch)


def test_structure_rules(tmp_path):
    from keff_cfd_b200 import structure_from_csv, structure_from_image
    img = np.array([[0, 149, 150], [255, 10, 200]], dtype=np.uint8)
    s = structure_from_image(img)
    assert s.tolist() == [[0, 0, 1], [1, 0, 1]]                    # pixel < 150 fluid (keff2D.cpp:113)
    s2 = structure_from_image(img, 2, 3)
    assert s2.shape == (6, 6) and all(s2[i, j] == s[i // 3, j // 2] for i in range(6) for j in range(6))
    p = tmp_path / "s.csv"
    p.write_text("1,2,3\n0,0,0\n3,1,2\n")
    v = structure_from_csv(str(p), 4, 4, 4)
    assert v.sum() == 3 and v[3, 2, 1] == 1 and v[0, 0, 0] == 1 and v[2, 1, 3] == 1   # s[z][y][x] (keff3D.h:275)
    g = gold_npy("img_00009.npy")
    assert abs(1 - structure_from_image(g).mean() - 0.3736572265625) < 1e-15           # reference porosity


def test_generators_are_reproducible():
    from keff_cfd_b200 import structures as S
    a = S.bernoulli_volume((6, 5, 8), 0.5, 1234)
    b = S.bernoulli_volume((3, 5, 8), 0.5, 1234, z0=3)
    assert np.array_equal(a[3:], b)                                # slabs see the global index
    assert int(S.splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF   # known splitmix64(0)
    assert abs(S.bernoulli_volume((32, 32, 32), 0.5, 7).mean() - 0.5) < 0.01
    i0, i1 = S.disc_image(3), S.disc_image(3)
    assert np.array_equal(i0, i1) and 0.25 < i0.mean() < 0.9
    assert S.to_gray(i0).max() == 255


def test_partition_helpers():
    from keff_cfd_b200 import dist as D
    for n, w in ((1024, 8), (10, 3), (7, 8), (10000, 8)):
        ext = [D.slab_extent(n, r, w) for r in range(w)]
        assert ext[0][0] == 0 and sum(e[1] for e in ext) == n
        assert all(ext[i][0] + ext[i][1] == ext[i + 1][0] for i in range(w - 1))
        assert max(e[1] for e in ext) - min(e[1] for e in ext) <= 1
