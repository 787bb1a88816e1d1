"""Two-level acceleration structure and MODE_UPDATE refits (RTAccelerationStructure::{createBLAS,
addInstance, buildTLAS}, rt_acceleration_structure.cpp:205-535, ALLOW_UPDATE :467 / MODE_UPDATE :508).

OHB_ACCEL_TWO_LEVEL builds one object-space BLAS per instance under a TLAS and maps rays into object space on descent;
OHB_ACCEL_FLATTEN (default) bakes the transforms into one world-space tree.  Both have a MODE_UPDATE path
(ohb_update_instances).  The oracle states the two-level arithmetic (object-space watertight test, the ray mapped with one
rounding per operation) so results are bit-exact per mode; the two modes agree with each other up to that rounding."""
import numpy as np
import pytest

from ohao_engine_b200 import scenes
from oracle import oracle_py as O
from tests import util


def instanced_scene(n_inst=40, seed=5, tris_per=300):
    """Blobs of very different sizes under general affine transforms (rotation + non-uniform scale + shear) over a ground quad."""
    rng = np.random.default_rng(seed)
    base = scenes.synthetic_2m(nblobs=1, tris_per_blob=tris_per, env_size=(64, 32))
    # one blob mesh in object space, reused with different transforms (each instance owns its copy of the triangles: one BLAS per actor)
    nt = base.ntris - 2
    P = base.positions[:, :3]; I = base.indices.reshape(-1, 3)[:nt]
    used = np.unique(I); remap = -np.ones(len(P), np.int64); remap[used] = np.arange(len(used))
    bp = P[used] - P[used].mean(0); bn = base.normals[used, :3]; bi = remap[I].astype(np.uint32).reshape(-1)
    meshes = []
    for k in range(n_inst):
        a = rng.normal(size=(3, 3)); q, _ = np.linalg.qr(a)
        sc = np.diag(rng.uniform(0.3, 3.0, 3) * (10.0 if k % 7 == 0 else 1.0)); sh = np.eye(3); sh[0, 1] = rng.uniform(-0.3, 0.3)
        m = np.zeros((3, 4), np.float32); m[:, :3] = (q @ sc @ sh).astype(np.float32); m[:, 3] = rng.uniform(-40, 40, 3) * np.array([1, 0.3, 1]) + np.array([0, 12, 0])
        me = scenes.Mesh(positions=bp.astype(np.float32), normals=bn.astype(np.float32), uvs=np.zeros((len(bp), 2), np.float32), indices=bi, xform=m.reshape(12),
                         base_color=tuple(rng.uniform(0.2, 0.9, 3)), roughness=float(rng.uniform(0.3, 1.0)), metallic=float(k % 3 == 0), name=f"blob{k}")
        meshes.append(me)
    g = 80.0
    ground = scenes.quad_mesh((-g, 0, -g), (-g, 0, g), (g, 0, g), (g, 0, -g), (0, 1, 0)); ground.base_color = (0.5, 0.5, 0.5); ground.roughness = 0.9
    meshes.append(ground)
    lights = [scenes.Light(position=(10.0, 40.0, 5.0), intensity=900.0, radius=2.0), scenes.Light(position=(-25.0, 30.0, -15.0), color=(1.0, 0.8, 0.6), intensity=500.0, radius=1.5)]
    return scenes.pack_scene(meshes, lights, env=scenes.procedural_env(64, 32), name="instanced")


def moved(ps, seed=9):
    rng = np.random.default_rng(seed)
    inst = ps.instances.copy()
    for k in range(len(inst) - 1):                       # every blob moves and spins; the ground stays
        m = inst[k]["xform"].reshape(3, 4).copy()
        a = rng.normal(size=(3, 3)); q, _ = np.linalg.qr(a)
        m[:, :3] = (q @ m[:, :3]).astype(np.float32); m[:, 3] += rng.uniform(-6, 6, 3).astype(np.float32) * np.array([1, 0.2, 1], np.float32)
        inst[k]["xform"] = m.reshape(12)
    return inst


@pytest.fixture(scope="module")
def inst_scene():
    return instanced_scene(), scenes.synthetic_camera()


def _check_rays(osc, other, rays, two_level):
    ref, got = osc.trace(rays), other.trace(rays)
    mism, ties = util.compare_hits(ref, got)
    assert mism == 0, (two_level, mism, ties)
    for k in ("t", "u", "v"):
        assert np.array_equal(ref[k], got[k]), (two_level, k)
    assert np.array_equal(osc.occluded(rays[:40000]) != 0, other.occluded(rays[:40000]) != 0)
    return ref


def _run_modes(ps, make_other, nrays):
    """Both modes against the oracle in the same mode, before and after a transform-only update; the modes against each other."""
    rays = util.random_rays(nrays, -60, 60, seed=31)
    hits = {}
    for two_level in (False, True):
        osc = O.OracleScene(ps); osc.set_accel_mode(two_level)
        other = make_other(two_level)
        a = _check_rays(osc, other, rays, two_level)
        assert (a["prim"] != 0xFFFFFFFF).mean() > 0.3
        new_inst = moved(ps)
        osc.update_instances(new_inst); other.update_instances(new_inst)
        b = _check_rays(osc, other, rays, two_level)
        assert (a["prim"] != b["prim"]).mean() > 0.05                                  # the update really changed what rays hit
        # a refit tree must give what a fresh build of the moved scene gives
        import copy
        ps2 = copy.deepcopy(ps); ps2.instances = new_inst
        fresh = O.OracleScene(ps2); fresh.set_accel_mode(two_level)
        c = fresh.trace(rays)
        for k in ("prim", "t", "u", "v"):
            assert np.array_equal(b[k], c[k]), k
        hits[two_level] = b
    same = hits[False]["prim"] == hits[True]["prim"]
    assert same.mean() > 0.9995                                                         # object-space vs world-space rounding flips a silhouette ray now and then
    ta, tb = hits[False]["t"][same], hits[True]["t"][same]                              # grazing hits amplify the rounding of the mapped ray
    assert np.isclose(ta, tb, rtol=1e-4, atol=1e-4).mean() > 0.999 and np.allclose(ta, tb, rtol=2e-2, atol=2e-2)


def test_two_level_and_refit_product_code_on_the_host(inst_scene):
    from tests.emul import emul_py as E
    ps, _ = inst_scene

    def make(two_level):
        e = E.EmulScene(ps); e.set_accel_mode(two_level); return e
    _run_modes(ps, make, 60000)


def test_two_level_render_matches_oracle_on_the_host(inst_scene):
    from tests.emul import emul_py as E
    ps, cam = inst_scene
    W, H, spp = 80, 45, 2
    osc = O.OracleScene(ps); osc.set_accel_mode(True)
    esc = E.EmulScene(ps); esc.set_accel_mode(True)
    ro = osc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True)
    re = esc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True)
    bad, worst = util.sample_parity(ro["samples"], re["samples"])
    assert bad < 2e-3 and worst < 2e-3, (bad, worst)
    assert ro["samples"][..., :3].max() > 0.05


def test_two_level_degenerate_instances():
    """Single-triangle and two-triangle instances (BLAS root = a leaf), one instance only (TLAS root = a leaf), a masked-out instance."""
    from tests.emul import emul_py as E
    tri = scenes.Mesh(positions=np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), normals=np.tile(np.array([0, 0, 1], np.float32), (3, 1)),
                      uvs=np.zeros((3, 2), np.float32), indices=np.arange(3, dtype=np.uint32), xform=scenes.trs(position=(0.2, 0.1, -1.0), scale=(2, 2, 2)))
    quad = scenes.quad_mesh((-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1), (0, 1, 0)); quad.xform = scenes.trs(position=(0, -1.5, 0), scale=(3, 1, 3))
    for meshes in ([tri], [tri, quad], [quad, tri, quad]):
        ps = scenes.pack_scene(meshes, [scenes.Light(position=(0, 3, 0))])
        if len(meshes) == 3:
            ps.instances["mask"][2] = 0; ps.instances["xform"][2][7] = 0.5                 # an invisible copy in front of the others
        rays = util.random_rays(6000, -2.5, 2.5, seed=len(meshes))
        osc = O.OracleScene(ps); osc.set_accel_mode(True)
        esc = E.EmulScene(ps); esc.set_accel_mode(True)
        ref = _check_rays(osc, esc, rays, True)
        assert (ref["prim"] != 0xFFFFFFFF).any()


@pytest.mark.gpu
def test_two_level_and_refit_gpu(inst_scene):
    from ohao_engine_b200 import binding as B
    ps, cam = inst_scene

    def make(two_level):
        r = B.Renderer(64, 64); r.set_accel_mode(two_level); r.set_scene(ps)
        st = r.accel_stats(); assert st.num_tris == ps.ntris and st.build_ms > 0
        return r
    _run_modes(ps, make, 400000)
    # render in two-level mode: per-sample radiance against the oracle in the same mode; update timing is reported
    W, H, spp = 320, 180, 2
    osc = O.OracleScene(ps); osc.set_accel_mode(True)
    r = B.Renderer(W, H); r.set_accel_mode(True); r.set_scene(ps)
    ro = osc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True)
    got = r.render(cam.view(), cam.proj(W, H), spp, dump=True)
    bad, worst = util.sample_parity(ro["samples"], got)
    assert bad < 3e-3 and worst < 2e-3, (bad, worst)
    st = r.update_instances(moved(ps))
    assert st.update_ms > 0
    print(f"\n[two-level] {len(ps.instances)} instances, {ps.ntris} tris: build {r.accel_stats().build_ms:.2f} ms, TLAS refit {st.update_ms:.3f} ms")
