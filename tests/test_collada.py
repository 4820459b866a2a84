"""CPU: the product's COLLADA loader (libmplb.so, host code) against the fixtures produced by the
independent numpy restatement, and its error behaviour (load_mesh.hpp:250-254)."""
import os

import numpy as np
import pytest

import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES
from conftest import GOLDEN, REFERENCE

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "resources")),
                               reason="reference resources only exist in the build container")


@needs_ref
@pytest.mark.parametrize("name", sorted(SE3_SCENES))
def test_loader_matches_fixture(name):
    want = np.load(os.path.join(GOLDEN, "scenes", name + ".npz"))
    sc = SE3_SCENES[name]
    env, _ = M.load_collada(os.path.join(REFERENCE, "resources", sc["env"]), False, False)
    robot, center = M.load_collada(os.path.join(REFERENCE, "resources", sc["robot"]), True, False)
    assert env.shape == want["env"].shape and robot.shape == want["robot"].shape
    np.testing.assert_allclose(env, want["env"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(robot, want["robot"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(center, want["center"], rtol=0, atol=1e-10)


@needs_ref
def test_identity_root_transform():
    want = np.load(os.path.join(GOLDEN, "scenes", "autolab.npz"))["env"]
    got, center = M.load_collada(os.path.join(REFERENCE, "resources", "AUTOLAB.dae"), False, True)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    assert (center == 0).all()
    rotated, _ = M.load_collada(os.path.join(REFERENCE, "resources", "AUTOLAB.dae"), False, False)
    # Z_UP root: (x,y,z) -> (x,z,-y)
    np.testing.assert_allclose(rotated[..., 0], got[..., 0], atol=1e-12)
    np.testing.assert_allclose(rotated[..., 1], got[..., 2], atol=1e-12)
    np.testing.assert_allclose(rotated[..., 2], -got[..., 1], atol=1e-12)


def test_missing_file_is_an_error():
    with pytest.raises(M.MplbError) as ei:
        M.load_collada("/no/such/mesh.dae")
    assert ei.value.code == -2 and "could not load mesh file" in str(ei.value)


def test_small_document(tmp_path):
    dae = tmp_path / "quad.dae"
    dae.write_text("""<?xml version="1.0"?>
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
 <asset><up_axis>Y_UP</up_axis></asset>
 <library_geometries><geometry id="g"><mesh>
  <source id="p"><float_array id="pa" count="12">0 0 0 1 0 0 1 1 0 0 1 0</float_array>
   <technique_common><accessor source="#pa" count="4" stride="3"/></technique_common></source>
  <vertices id="v"><input semantic="POSITION" source="#p"/></vertices>
  <polylist count="1"><input semantic="VERTEX" source="#v" offset="0"/><vcount>4</vcount><p>0 1 2 3</p></polylist>
  <lines count="1"><input semantic="VERTEX" source="#v" offset="0"/><p>0 1</p></lines>
 </mesh></geometry></library_geometries>
 <library_visual_scenes><visual_scene id="s"><node>
   <matrix>1 0 0 10  0 1 0 20  0 0 1 30  0 0 0 1</matrix>
   <node><matrix>2 0 0 0  0 2 0 0  0 0 2 0  0 0 0 1</matrix><instance_geometry url="#g"/></node>
 </node></visual_scene></library_visual_scenes>
 <scene><instance_visual_scene url="#s"/></scene>
</COLLADA>""")
    tris, _ = M.load_collada(str(dae))
    want = np.array([[[0, 0, 0], [1, 0, 0], [1, 1, 0]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]]], dtype=float) * 2 + [10, 20, 30]
    np.testing.assert_allclose(tris, want)
    shifted, center = M.load_collada(str(dae), True)
    # 4 quad corners + 2 line vertices (a separate mesh, counted again like the importer does)
    np.testing.assert_allclose(center, (np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 0], [1, 0, 0]]).mean(0) * 2 + [10, 20, 30]))
    np.testing.assert_allclose(shifted + center, want)
