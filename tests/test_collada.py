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


SYNTHETIC = """<?xml version="1.0"?>
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
 <asset><unit meter="1"/><up_axis>Z_UP</up_axis></asset>
 <library_geometries>
  <geometry id="box"><mesh>
   <source id="bp"><float_array id="bpa" count="24">-1 -1 -1  1 -1 -1  1 1 -1  -1 1 -1  -1 -1 1  1 -1 1  1 1 1  -1 1 1</float_array>
    <technique_common><accessor source="#bpa" count="8" stride="3"/></technique_common></source>
   <source id="bn"><float_array id="bna" count="6">0 0 1  0 0 -1</float_array>
    <technique_common><accessor source="#bna" count="2" stride="3"/></technique_common></source>
   <vertices id="bv"><input semantic="POSITION" source="#bp"/></vertices>
   <triangles count="4"><input semantic="VERTEX" source="#bv" offset="0"/><input semantic="NORMAL" source="#bn" offset="1"/>
    <p>0 1 1 1 2 1   0 1 2 1 3 1   4 0 6 0 5 0   4 0 7 0 6 0</p></triangles>
   <polylist count="2"><input semantic="VERTEX" source="#bv" offset="0"/><input semantic="NORMAL" source="#bn" offset="1"/>
    <vcount>4 5</vcount><p>0 0 1 0 5 0 4 0   1 1 2 1 6 1 5 1 3 0</p></polylist>
   <lines count="2"><input semantic="VERTEX" source="#bv" offset="0"/><p>0 6 1 7</p></lines>
  </mesh></geometry>
  <geometry id="tri"><mesh>
   <source id="tp"><float_array id="tpa" count="9">0.5 0.25 0  3 0 0.125  0 2 7</float_array>
    <technique_common><accessor source="#tpa" count="3" stride="3"/></technique_common></source>
   <vertices id="tv"><input semantic="POSITION" source="#tp"/></vertices>
   <triangles count="2"><input semantic="VERTEX" source="#tv" offset="0"/><p>0 1 2 2 1 0</p></triangles>
  </mesh></geometry>
 </library_geometries>
 <library_visual_scenes><visual_scene id="s">
  <node><matrix>0 -1 0 4  1 0 0 -3  0 0 1 0.5  0 0 0 1</matrix><instance_geometry url="#box"/>
   <node><matrix>1.5 0 0 0  0 0.5 0 2  0 0 3 0  0 0 0 1</matrix><instance_geometry url="#tri"/><instance_geometry url="#box"/></node>
  </node>
  <node><matrix>1 0 0 100  0 1 0 0  0 0 1 0  0 0 0 1</matrix><instance_geometry url="#tri"/></node>
 </visual_scene></library_visual_scenes>
 <scene><instance_visual_scene url="#s"/></scene>
</COLLADA>"""


@pytest.mark.parametrize("shift,identity", [(False, False), (True, False), (False, True), (True, True)])
def test_synthetic_document_matches_the_numpy_restatement(tmp_path, shift, identity):
    """Runs anywhere (no reference resources): a document with nested node transforms, Z_UP, two geometries instanced
    three times, triangles / polylists with 4- and 5-gons / lines, shared and repeated corners -- the C++ importer of
    libmplb.so against the independent numpy/ElementTree restatement (oracle/collada_np.py)."""
    from oracle import collada_np
    dae = tmp_path / "synthetic.dae"
    dae.write_text(SYNTHETIC)
    got, center = M.load_collada(str(dae), shift, identity)
    want, wcenter = collada_np.load_collada(str(dae), shift, identity, return_center=True)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    np.testing.assert_allclose(center, wcenter, rtol=0, atol=1e-12)
    assert len(got) == 2 * (4 + 2 + 3) + 2 * 2    # two box instances (4 tris + quad + pentagon), two triangle-pair instances


def test_disk_cache_hits_misses_and_damaged_entries(tmp_path):
    """mplb_set_cache_dir: the second import of the same bytes is a hit with identical output; other flags or other
    bytes are different entries; a truncated or bit-flipped entry is ignored and rewritten."""
    import ctypes as C
    from mplambda_b200 import _capi
    L = _capi.lib()
    dae = tmp_path / "synthetic.dae"
    dae.write_text(SYNTHETIC)
    cache = tmp_path / "cache"

    def counts():
        h, m = C.c_uint64(), C.c_uint64()
        L.mplb_cache_stats(C.byref(h), C.byref(m))
        return h.value, m.value
    try:
        assert L.mplb_set_cache_dir(str(cache).encode()) == 0
        h0, m0 = counts()
        first, c1 = M.load_collada(str(dae), True, False)
        assert counts() == (h0, m0 + 1) and len(list(cache.iterdir())) == 1
        again, c2 = M.load_collada(str(dae), True, False)
        assert counts() == (h0 + 1, m0 + 1)
        np.testing.assert_array_equal(first, again)
        np.testing.assert_array_equal(c1, c2)
        M.load_collada(str(dae), False, False)                     # other flags: its own entry
        assert counts() == (h0 + 1, m0 + 2) and len(list(cache.iterdir())) == 2
        entry = sorted(cache.iterdir(), key=lambda p: p.stat().st_mtime)[0]
        raw = bytearray(entry.read_bytes())
        raw[-5] ^= 0x40                                            # damaged payload: checksum mismatch
        entry.write_bytes(bytes(raw))
        third, _ = M.load_collada(str(dae), True, False)
        assert counts() == (h0 + 1, m0 + 3)
        np.testing.assert_array_equal(first, third)
        entry.write_bytes(bytes(raw[:40]))                         # truncated
        M.load_collada(str(dae), True, False)
        assert counts() == (h0 + 1, m0 + 4)
        M.load_collada(str(dae), True, False)
        assert counts() == (h0 + 2, m0 + 4)                        # rewritten entry is good again
        with pytest.raises(M.MplbError):                           # errors are not cached
            M.load_collada(str(tmp_path / "missing.dae"))
        blocked = tmp_path / "a_file"
        blocked.write_text("x")
        assert L.mplb_set_cache_dir(str(blocked).encode()) != 0
    finally:
        L.mplb_set_cache_dir(None)
