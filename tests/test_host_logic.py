"""Host-side logic: mesh readers/generators and the case-file format (CPU)."""
import numpy as np

from gpupfem2_b200 import casefile
from gpupfem2_b200.mesh import load_dat, poiseuille_field, structured_channel


def test_load_dat_format(tmp_path):
    # reference format (src/mesh_2d.cu:36-96): header, vertices "id x y z", entities "id type ...", 203 = triangle
    p = tmp_path / "m.dat"
    p.write_text("4 4\n1 0.0 0.0 0.0\n2 1.0 0.0 0.0\n3 1.0 1.0 0.0\n4 0.0 1.0 0.0\n"
                 "1 102 1 2 \n2 102 2 3 \n3 203 1 2 3 \n4 203 1 3 4 \n")
    m = load_dat(str(p), scale=2.0)
    assert m.n_nodes == 4 and m.n_cells == 2
    assert np.array_equal(m.cells, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32))  # 1-based -> 0-based
    assert np.array_equal(m.vertices[2], [2.0, 2.0])


def test_structured_channel_numbering():
    for colmajor in (True, False):
        m = structured_channel(5, 3, 10.0, 3.0, colmajor=colmajor)
        assert m.n_nodes == 6 * 4 and m.n_cells == 2 * 5 * 3
        v, c = m.vertices, m.cells.astype(np.int64)
        a, b, d = v[c[:, 0]], v[c[:, 1]], v[c[:, 2]]
        area = 0.5 * ((b[:, 0] - a[:, 0]) * (d[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (d[:, 0] - a[:, 0]))
        assert np.allclose(area, 0.5 * 2.0 * 1.0) and np.all(area > 0)  # counter-clockwise
        assert np.isclose(area.sum(), 30.0)
        if colmajor:  # a strip in x is a contiguous cell range: centroid x is non-decreasing in blocks of 2*ny
            cx = (a[:, 0] + b[:, 0] + d[:, 0]) / 3
            blocks = cx.reshape(5, 6)
            assert np.all(blocks.max(axis=1)[:-1] < blocks.min(axis=1)[1:])


def test_poiseuille_field():
    m = structured_channel(2, 4, 1.0, 2.0)
    fx, fy = poiseuille_field(m, 3.0, 2.0)
    assert np.all(fy == 0)
    assert np.isclose(fx.max(), 3.0) and fx.min() == 0.0


def test_casefile_roundtrip(tmp_path, oracle):
    m = oracle.complete_mesh(structured_channel(3, 2, 1.0, 1.0))
    fx, fy = poiseuille_field(m, 1.0, 1.0)
    p = tmp_path / "c.bin"
    casefile.write_case(str(p), m, fx, fy, 2, 3, 0.01, 5, (1, 5))
    raw = np.fromfile(str(p), dtype=np.uint8)
    hdr = raw[:64].view(np.int64)
    assert hdr[0] == casefile.MAGIC and hdr[1] == m.n_nodes and hdr[2] == m.n_cells and hdr[7] == 2
    expect = 64 + 8 + 16 + m.n_nodes * 16 + m.n_cells * 12 + (m.n_cells + 1) * 4 + m.nbr_indices.size * 4 + m.n_nodes * 16
    assert raw.size == expect
