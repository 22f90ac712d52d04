"""GPU parity tests (run with `-m gpu` on the B200 box).  They call the product path through the C ABI
(include/mgs_rasterizer.h via manigaussian_b200.rasterizer) and compare it with
  (1) the CPU oracle oracle/gs_oracle.c on the same seeded inputs, stage by stage, and
  (2) the compiled, unmodified reference rasterizer in oracle/_ref when that build travelled with the snapshot.
Bars (BASELINE.json north_star): tile ids / sort keys / ranges / radii bit-exact; images and all gradient tensors
within 1e-4 relative L2.
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

TOL = 1e-4  # relative L2, stated by BASELINE.json's north_star (the bar against the reference rasterizer)
# Against the CPU oracle the same bar holds except where fp32 evaluation order dominates: `ragged_f32` has screen-filling
# splats (thousands of signed terms per pixel, cancelling sums), where gcc's uncontracted fp32 and nvcc's FMA-contracted
# fp32 differ by 2-5e-4 on gradients -- the compiled reference shows the SAME distance to the oracle (ref_vs_oracle in
# profiles/r1_parity_report.json) while ours-vs-reference stays at 1e-6.
ORACLE_GRAD_TOL = {"ragged_f32": 1e-3}

CASES = {
    "tiny_f3": dict(P=300, W=32, H=32, F=3, seed=1),
    "mg_real_f3": dict(P=16384, W=128, H=128, F=3, seed=2),                   # ManiGaussian's real regime
    "ragged_f32": dict(P=5000, W=200, H=120, F=32, seed=3, bg=(0.3, 0.6, 0.9)),
    "rgb_only_precomp": dict(P=4000, W=96, H=96, F=0, seed=4, precomp_colors=True, precomp_cov=True),
    "sh3_big": dict(P=1500, W=128, H=96, F=8, seed=5, sh_degree=3, scale0=0.04, bg=(1.0, 1.0, 1.0)),
    "c1_like": dict(P=50000, W=128, H=128, F=0, seed=6),
    "f16": dict(P=3000, W=64, H=64, F=16, seed=7),
    "f5_odd": dict(P=2000, W=64, H=80, F=5, seed=8),
}


def _live(fw):
    return fw["radii"] > 0


def check_stagewise_vs_oracle(inp, grad_tol=TOL):
    """Ours vs the C oracle.  Integer stages are checked exactly by feeding OUR upstream outputs to the oracle's
    downstream stage (so the FMA-contraction difference of the projection cannot leak into an index compare)."""
    import ctypes as C
    from oracle import gs_oracle as O
    ours, ours_bw = util.run_ours(inp)
    orc, orc_bw = util.run_oracle(inp)
    P, W, H, F = inp["P"], inp["W"], inp["H"], inp["F"]
    # -- projection: floats by tolerance, integer outputs identical except for documented borderline cases
    live = _live(ours) & _live(orc)
    assert (_live(ours) != _live(orc)).mean() <= 1e-4
    assert (ours["radii"] != orc["radii"]).mean() <= 1e-4
    assert (ours["tiles_touched"] != orc["tiles_touched"]).mean() <= 1e-4
    for k in ("depths", "means2D", "conic_opacity", "cov3D"):
        if k == "cov3D" and inp["g"]["cov3D_precomp"] is not None:
            continue  # not computed when the caller supplies it
        assert util.rel_l2(ours[k][live], orc[k][live]) < 1e-5, k
    if inp["g"]["colors_precomp"] is None:
        assert util.rel_l2(ours["rgb"][live], orc["rgb"][live]) < 1e-5
        assert (ours["clamped"][live] != orc["clamped"][live]).mean() <= 1e-4
    # -- binning: bit-exact given our own projection outputs
    L = O.lib()
    R = ours["num_rendered"]
    off = np.zeros(P, np.uint32)
    assert int(L.gso_scan(C.c_int(P), O._p(ours["tiles_touched"]), O._p(off))) == R
    assert int(ours["point_offsets"].astype(np.uint32)[-1]) == R  # ours scans in depth order: only the total is comparable
    ku, vu = np.zeros(R, np.uint64), np.zeros(R, np.uint32)
    m2, dp, rd = np.ascontiguousarray(ours["means2D"]), np.ascontiguousarray(ours["depths"]), np.ascontiguousarray(ours["radii"])
    dp = np.where(rd > 0, dp, 0).astype(np.float32)  # culled Gaussians carry +inf depth in our state
    L.gso_duplicate_with_keys(C.c_int(P), O._p(m2), O._p(dp), O._p(off), O._p(rd), C.c_int(W), C.c_int(H), O._p(ku), O._p(vu))
    # our emission is in depth order (then one stable per-tile pass): same multiset of instances as the reference's
    assert np.array_equal(np.sort(vu), np.sort(ours["point_list_unsorted"]))
    live_order = ours["depth_order"][ours["radii"][ours["depth_order"]] > 0]
    assert np.all(np.diff(ours["depths"].view(np.uint32)[live_order].astype(np.int64)) >= 0), "depth order not sorted"
    ks, vs = np.zeros(R, np.uint64), np.zeros(R, np.uint32)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    L.gso_sort_pairs(C.c_uint32(R), O._p(ku), O._p(vu), O._p(ks), O._p(vs), C.c_int(32 + O.get_higher_msb(T)))
    assert np.array_equal(ks, ours["point_list_keys"]), "sorted keys differ"
    assert np.array_equal(vs, ours["point_list"]), "sorted values differ"
    rg = np.zeros((T, 2), np.uint32)
    L.gso_identify_tile_ranges(C.c_uint32(R), O._p(ks), C.c_int(T), O._p(rg))
    assert np.array_equal(rg, ours["ranges"])
    # -- forward blend: oracle on OUR sorted list and projected records
    colors = inp["g"]["colors_precomp"] if inp["g"]["colors_precomp"] is not None else np.ascontiguousarray(ours["rgb"])
    N = W * H
    fT, nc = np.zeros(N, np.float32), np.zeros(N, np.uint32)
    oc = np.zeros((3, H, W), np.float32)
    of = np.zeros((max(F, 1), H, W), np.float32)
    co = np.ascontiguousarray(ours["conic_opacity"])
    L.gso_render_forward(C.c_int(W), C.c_int(H), C.c_int(F), O._p(rg), O._p(vs), O._p(m2), O._p(colors),
                         O._p(inp["g"]["feature"]) if F else C.c_void_p(0), O._p(co), O._p(inp["bg"]), O._p(fT), O._p(nc),
                         O._p(oc), O._p(of))
    assert util.rel_l2(ours["out_color"], oc) < 1e-5
    if F:
        assert util.rel_l2(ours["out_feature"], of[:F]) < 1e-5
    assert util.rel_l2(ours["final_T"], fT) < 1e-5
    assert (ours["n_contrib"] != nc).mean() <= 1e-4
    # -- whole pipeline, end to end, against the oracle's own run
    assert util.rel_l2(ours["out_color"], orc["out_color"]) < TOL
    if F:
        assert util.rel_l2(ours["out_feature"], orc["out_feature"]) < TOL
    for k in ours_bw:
        if k == "dL_dfeature" and not F:
            continue
        assert util.rel_l2(ours_bw[k], orc_bw[k]) < grad_tol, (k, util.rel_l2(ours_bw[k], orc_bw[k]))
    return ours, ours_bw


@pytest.mark.parametrize("name", list(CASES))
def test_vs_oracle(name):
    check_stagewise_vs_oracle(util.make_inputs(**CASES[name]), ORACLE_GRAD_TOL.get(name, TOL))


@pytest.mark.parametrize("name", list(CASES))
def test_vs_compiled_reference(name):
    inp = util.make_inputs(**CASES[name])
    ref, ref_bw = util.run_reference(inp)
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources are only available in the build container)")
    ours, ours_bw = util.run_ours(inp)
    F = inp["F"]
    # bit-exact: everything that decides tile ids and sort keys
    assert np.array_equal(ours["radii"], ref["radii"])
    assert np.array_equal(ours["tiles_touched"], ref["tiles_touched"])
    live = _live(ref)
    for k in ("depths", "means2D", "conic_opacity"):
        assert np.array_equal(ours[k][live].view(np.uint32), ref[k][live].view(np.uint32)), k + " not bit-exact"
    assert ours["num_rendered"] == ref["num_rendered"]
    assert np.array_equal(ours["point_list_keys"], ref["point_list_keys"]), "sort keys differ from the reference"
    assert np.array_equal(ours["point_list"], ref["point_list"]), "sorted Gaussian ids differ from the reference"
    assert np.array_equal(ours["ranges"], ref["ranges"])
    assert (ours["n_contrib"] != ref["n_contrib"]).mean() <= 1e-5
    # 1e-4 relative L2: images, per-pixel state, every gradient tensor
    assert util.rel_l2(ours["out_color"], ref["out_color"]) < TOL
    if F:
        assert util.rel_l2(ours["out_feature"], ref["out_feature"]) < TOL
    assert util.rel_l2(ours["final_T"], ref["final_T"]) < TOL
    for k in ours_bw:
        if k == "dL_dfeature" and not F:
            continue
        if k == "dL_dcolors" and inp["g"]["colors_precomp"] is None:
            pass  # internal gradient in the SH path; still comparable
        assert util.rel_l2(ours_bw[k], ref_bw[k]) < TOL, (k, util.rel_l2(ours_bw[k], ref_bw[k]))


def test_depth_plane_matches_feature_channel():
    """The reference renders no depth; parity for the depth plane is obtained by feeding view-space z through a
    feature channel of the oracle (SURVEY.md finding 4) and comparing both the plane and dL/dmeans3D."""
    import torch
    from manigaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer
    inp = util.make_inputs(P=3000, W=64, H=64, F=0, seed=21)
    cam, g = inp["cam"], inp["g"]
    dev = "cuda"
    t = lambda x: torch.from_numpy(x).to(dev)
    means = t(g["means3D"]).requires_grad_(True)
    st = GaussianRasterizationSettings(64, 64, cam["tanfovx"], cam["tanfovy"], t(inp["bg"]), 1.0, t(cam["viewmatrix"]),
                                       t(cam["projmatrix"]), 1, t(cam["campos"]), False, False, False)
    rast = GaussianRasterizer(st, return_depth=True)
    color, feat, radii, depth = rast(means3D=means, means2D=torch.zeros_like(means), opacities=t(g["opacities"]), shs=t(g["shs"]),
                                     scales=t(g["scales"]), rotations=t(g["rotations"]))
    gd = torch.from_numpy(np.random.default_rng(0).normal(size=(64, 64)).astype(np.float32)).to(dev)
    (depth * gd).sum().backward()
    # oracle: z as a 1-channel feature; chain dL/dz to the means by hand (z = V[2,:3] . p + V[2,3])
    from oracle import gs_oracle as O
    vm = cam["viewmatrix"].reshape(-1)
    z = (g["means3D"] @ np.array([vm[2], vm[6], vm[10]], np.float32) + vm[14]).astype(np.float32)[:, None]
    kw = dict(scales=g["scales"], rotations=g["rotations"], shs=g["shs"], sh_degree=1, feature=z)
    fw = O.forward(g["means3D"], g["opacities"], cam["viewmatrix"], cam["projmatrix"], cam["campos"], 64, 64, cam["tanfovx"],
                   cam["tanfovy"], inp["bg"], **kw)
    assert util.rel_l2(depth.detach().cpu().numpy(), fw["out_feature"][0]) < TOL
    bw = O.backward(fw, np.zeros((3, 64, 64), np.float32), gd.cpu().numpy()[None], g["means3D"], cam["viewmatrix"], cam["projmatrix"],
                    cam["campos"], cam["tanfovx"], cam["tanfovy"], inp["bg"], **kw)
    dmean = bw["dL_dmeans3D"] + bw["dL_dfeature"] * np.array([vm[2], vm[6], vm[10]], np.float32)[None, :]
    assert util.rel_l2(means.grad.cpu().numpy(), dmean) < TOL


def test_autograd_module_like_manigaussian_render():
    """Drives the module exactly as agents/manigaussian_bc/gaussian_renderer/__init__.py:54-84 does (SH degree 1,
    L2-normalised 3-channel features, zero means2D with requires_grad) and checks gradients reach every input."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    inp = util.make_inputs(P=16384, W=128, H=128, F=3, seed=31)
    cam, g = inp["cam"], inp["g"]
    t = lambda x: torch.from_numpy(x).cuda()
    xyz, rot, scale = t(g["means3D"]).requires_grad_(True), t(g["rotations"]).requires_grad_(True), t(g["scales"]).requires_grad_(True)
    opa, shs, feat = t(g["opacities"]).requires_grad_(True), t(g["shs"]).requires_grad_(True), t(g["feature"]).requires_grad_(True)
    screenspace_points = torch.zeros_like(xyz, dtype=torch.float32, requires_grad=True, device="cuda") + 0
    screenspace_points.retain_grad()
    st = GaussianRasterizationSettings(image_height=128, image_width=128, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
                                       bg=torch.tensor([0, 0, 0], dtype=torch.float32, device="cuda"), scale_modifier=1.0,
                                       viewmatrix=t(cam["viewmatrix"]), projmatrix=t(cam["projmatrix"]), sh_degree=1,
                                       campos=t(cam["campos"]), prefiltered=False, debug=False, include_feature=True)
    lf = feat / (feat.norm(dim=-1, keepdim=True) + 1e-12)
    img, emb, radii = GaussianRasterizer(raster_settings=st)(means3D=xyz, means2D=screenspace_points, shs=shs, colors_precomp=None,
                                                            language_feature_precomp=lf, opacities=opa, scales=scale,
                                                            rotations=rot, cov3D_precomp=None)
    assert img.shape == (3, 128, 128) and emb.shape == (3, 128, 128) and radii.dtype == torch.int32
    loss = ((img - 0.5) ** 2).mean() + (1 - torch.nn.functional.cosine_similarity(emb, torch.ones_like(emb), dim=0)).mean()
    loss.backward()
    for v in (xyz, rot, scale, opa, shs, feat, screenspace_points):
        assert v.grad is not None and torch.isfinite(v.grad).all() and v.grad.abs().sum() > 0
    assert screenspace_points.grad.shape == (16384, 3) and (screenspace_points.grad[:, 2] == 0).all()


def test_errors_and_edge_cases():
    import torch
    from manigaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer
    inp = util.make_inputs(P=64, W=32, H=32, F=0, seed=41)
    cam, g = inp["cam"], inp["g"]
    t = lambda x: torch.from_numpy(x).cuda()
    st = GaussianRasterizationSettings(32, 32, cam["tanfovx"], cam["tanfovy"], t(np.array([0.1, 0.2, 0.3], np.float32)), 1.0,
                                       t(cam["viewmatrix"]), t(cam["projmatrix"]), 1, t(cam["campos"]), False, True, False)
    rast = GaussianRasterizer(st)
    m = t(g["means3D"])
    with pytest.raises(Exception):  # neither SHs nor colours (reference __init__.py:201-202)
        rast(means3D=m, means2D=m, opacities=t(g["opacities"]), scales=t(g["scales"]), rotations=t(g["rotations"]))
    with pytest.raises(Exception):  # both scale/rotation and cov3D (reference __init__.py:204-205)
        rast(means3D=m, means2D=m, opacities=t(g["opacities"]), shs=t(g["shs"]), scales=t(g["scales"]), rotations=t(g["rotations"]),
             cov3D_precomp=torch.zeros(64, 6).cuda())
    # all Gaussians behind the camera -> image == background, radii == 0, gradients exactly zero
    behind = (m * 0 + torch.tensor([10.0, 0.0, 1.1]).cuda()).requires_grad_(True)
    color, feat, radii = rast(means3D=behind, means2D=torch.zeros_like(behind), opacities=t(g["opacities"]), shs=t(g["shs"]),
                              scales=t(g["scales"]), rotations=t(g["rotations"]))
    assert (radii == 0).all() and feat.shape == (1,)
    assert torch.allclose(color, t(np.array([0.1, 0.2, 0.3], np.float32))[:, None, None].expand(3, 32, 32))
    color.sum().backward()
    assert (behind.grad == 0).all()
    # P == 0 (rasterize_points.cu:92,186)
    e = torch.zeros((0, 3)).cuda()
    color0, _, radii0 = rast(means3D=e, means2D=e, opacities=torch.zeros((0, 1)).cuda(), shs=torch.zeros((0, 4, 3)).cuda(),
                             scales=e, rotations=torch.zeros((0, 4)).cuda())
    assert color0.shape == (3, 32, 32) and (color0 == 0).all() and radii0.numel() == 0
    # markVisible == near-plane test (rasterizer_impl.cu:54-66)
    vis = rast.markVisible(m)
    from oracle import gs_oracle as O
    assert np.array_equal(vis.cpu().numpy(), O.mark_visible(g["means3D"], cam["viewmatrix"], cam["projmatrix"]))


def test_full_size_properties_c3():
    """BASELINE.json configs[2] size (500k Gaussians, 256x256, 32 feature channels): size-independent properties."""
    inp = util.make_inputs(P=500000, W=256, H=256, F=32, seed=51)
    ours, bw = util.run_ours(inp)
    keys, vals, R = ours["point_list_keys"], ours["point_list"], ours["num_rendered"]
    assert R == int(ours["tiles_touched"].sum()) == int(ours["point_offsets"][-1])
    assert np.all(keys[1:] >= keys[:-1]), "keys not sorted"
    same = keys[1:] == keys[:-1]
    assert np.all(vals[1:][same] > vals[:-1][same]), "stable order broken among equal keys"
    assert np.array_equal(np.sort(vals), np.sort(ours["point_list_unsorted"]))
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    rg = ours["ranges"].astype(np.int64)
    cnt = np.bincount(tiles, minlength=rg.shape[0])
    assert np.array_equal(rg[:, 1] - rg[:, 0], cnt)
    assert np.array_equal(keys.astype(np.uint32), ours["depths"][vals].view(np.uint32)), "low key bits != depth bits"
    assert np.all((ours["final_T"] >= 0) & (ours["final_T"] <= 1))
    assert np.all(ours["n_contrib"] <= (rg[:, 1] - rg[:, 0]).reshape(16, 16).repeat(16, 0).repeat(16, 1).ravel())
    dead = ours["radii"] == 0
    for k, v in bw.items():
        assert np.isfinite(v).all(), k
        if v.shape[0] == inp["P"]:
            assert not np.any(v[dead]), k + ": culled Gaussians must have exactly zero gradient"
    # linearity of the backward in the cotangent: grads(2g) == 2 grads(g) up to atomics' summation order
    inp2 = dict(inp)
    inp2["ct"] = {k: (None if v is None else 2 * v) for k, v in inp["ct"].items()}
    _, bw2 = util.run_ours(inp2)
    for k in ("dL_dmeans3D", "dL_dfeature", "dL_dopacity", "dL_dscales"):
        assert util.rel_l2(bw2[k], 2 * bw[k]) < 1e-5, k


def test_multi_view_api_and_accumulate_mode():
    """rasterize_views_raw (one C call, one stream per view, no host synchronisation, binning state sized from the instance
    counts of earlier calls) must reproduce the single-view entry point bit for bit, and the multi-view backward (sum over
    views in registers) must equal the sum of the per-view gradients (sum over views on one GPU == what the
    all-reduce produces across GPUs).  Also: an undersized binning state is reported, not fatal."""
    import torch
    from manigaussian_b200 import rasterizer as R
    from manigaussian_b200 import GaussianRasterizationSettings as S
    from manigaussian_b200.parallel import PackedGradients
    V, P, W, H, F = 3, 20000, 96, 80, 32
    inps = [util.make_inputs(P=P, W=W, H=H, F=F, seed=77, view=v, num_views=V) for v in range(V)]
    g = inps[0]["g"]
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    G = {k: t(g[k]) for k in ("means3D", "scales", "rotations", "opacities", "shs", "feature")}
    e = torch.Tensor([])
    views = [S(H, W, i["cam"]["tanfovx"], i["cam"]["tanfovy"], t(i["bg"]), 1.0, t(i["cam"]["viewmatrix"]), t(i["cam"]["projmatrix"]), 1,
               t(i["cam"]["campos"]), False, False, True) for i in inps]
    cts = [(t(i["ct"]["dL_dcolor"]), t(i["ct"]["dL_dfeature"])) for i in inps]
    R.reset_capacity_estimates()
    for rep in range(2):  # first call calibrates the capacities with a host read, the second runs on the history alone
        outs, sts = R.rasterize_views_raw(views, G["means3D"], e, G["feature"], G["opacities"], G["scales"], G["rotations"], 1.0, e, G["shs"], 1, True)
    pk = PackedGradients(P, F, 4, "cuda", zero=False)
    m2d = torch.empty((V, P, 3), device="cuda")
    R.rasterize_views_backward_raw(views, outs, sts, [c[0] for c in cts], [c[1] for c in cts], G["means3D"], e, G["feature"], G["scales"],
                                   G["rotations"], 1.0, e, G["shs"], 1, True, accumulate_into=pk.views, means2D_per_view=m2d)
    summed = R.rasterize_views_backward_raw(views, outs, sts, [c[0] for c in cts], [c[1] for c in cts], G["means3D"], e, G["feature"],
                                            G["scales"], G["rotations"], 1.0, e, G["shs"], 1, True)
    torch.cuda.synchronize()
    names = ("dL_dmeans2D", "dL_dcolors", "dL_dfeature", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations")
    total = {}
    for v, s in enumerate(views):
        single = R.rasterize_gaussians_raw(s.bg, G["means3D"], e, G["feature"], G["opacities"], G["scales"], G["rotations"], 1.0, e, s.viewmatrix,
                                           s.projmatrix, s.tanfovx, s.tanfovy, H, W, G["shs"], 1, s.campos, False, False, True)
        assert single[0] == outs[v].num_rendered() and not outs[v].overflowed() and outs[v][0] >= single[0]
        assert torch.equal(single[1], outs[v][1]) and torch.equal(single[2], outs[v][2]) and torch.equal(single[3], outs[v][3])
        gr = R.rasterize_gaussians_backward_raw(s.bg, G["means3D"], single[3], e, G["feature"], G["scales"], G["rotations"], 1.0, e, s.viewmatrix,
                                                s.projmatrix, s.tanfovx, s.tanfovy, cts[v][0], cts[v][1], G["shs"], 1, s.campos, single[4],
                                                single[0], single[5], single[6], False, True)
        assert util.rel_l2(m2d[v].cpu().numpy(), gr[0].cpu().numpy()) < 1e-5
        for n, x in zip(names, gr):
            total[n] = x.double() if n not in total else total[n] + x.double()
    for k, view in pk.views.items():
        assert util.rel_l2(view.cpu().numpy(), total[k].reshape(view.shape).cpu().numpy()) < 1e-5, k
    for n, x in zip(names, summed):
        if x is not None and n in total and n not in ("dL_dcolors", "dL_dcov3D"):
            assert util.rel_l2(x.cpu().numpy(), total[n].reshape(x.shape).cpu().numpy()) < 1e-5, n
    # an undersized binning state: the overflow is flagged and the farthest instances are dropped, nothing else breaks
    small = [max(1024, outs[v].num_rendered() // 2) for v in range(V)]
    with torch.no_grad():
        o2, _ = R.rasterize_views_raw(views, G["means3D"], e, G["feature"], G["opacities"], G["scales"], G["rotations"], 1.0, e, G["shs"], 1, True,
                                      capacities=small)
    assert all(o.overflowed() for o in o2) and all(torch.isfinite(o[1]).all() for o in o2)


def test_dynamic_path_gradients_reach_deformation_offsets():
    """BASELINE.json configs[3] / SURVEY.md 3.5: the 'dyna' render uses mu + d_mu, normalize(r + d_r) and (extension)
    s + d_s (models_embed.py:297-304); the offsets' gradients are the rasterizer's dL/dmeans3D, dL/drotations chained
    through normalize, and dL/dscales.  Checked against the oracle with the chain rule done in float64 on the CPU."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    inp = util.make_inputs(P=6000, W=96, H=96, F=32, seed=91, scale0=0.012)  # moderate splats: see ORACLE_GRAD_TOL
    cam, g, ct = inp["cam"], inp["g"], inp["ct"]
    rng = np.random.default_rng(5)
    d_mu = rng.normal(0, 0.01, g["means3D"].shape).astype(np.float32)
    d_r = rng.normal(0, 0.05, g["rotations"].shape).astype(np.float32)
    d_s = (rng.normal(0, 0.1, g["scales"].shape) * g["scales"]).astype(np.float32)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    tmu, tr, ts = t(d_mu).requires_grad_(True), t(d_r).requires_grad_(True), t(d_s).requires_grad_(True)
    means = t(g["means3D"]) + tmu
    rots = torch.nn.functional.normalize(t(g["rotations"]) + tr, dim=-1)
    scales = t(g["scales"]) + ts
    st = GaussianRasterizationSettings(96, 96, cam["tanfovx"], cam["tanfovy"], t(inp["bg"]), 1.0, t(cam["viewmatrix"]),
                                       t(cam["projmatrix"]), 1, t(cam["campos"]), False, False, True)
    img, emb, _ = GaussianRasterizer(st)(means3D=means, means2D=torch.zeros_like(means), opacities=t(g["opacities"]), shs=t(g["shs"]),
                                         language_feature_precomp=t(g["feature"]), scales=scales, rotations=rots)
    ((img * t(ct["dL_dcolor"])).sum() + (emb * t(ct["dL_dfeature"])).sum()).backward()
    # oracle on the deformed cloud
    g2 = dict(g)
    q = g["rotations"].astype(np.float64) + d_r
    g2["means3D"] = (g["means3D"] + d_mu).astype(np.float32)
    g2["rotations"] = rots.detach().cpu().numpy()
    g2["scales"] = (g["scales"] + d_s).astype(np.float32)
    inp2 = dict(inp)
    inp2["g"] = g2
    # expected gradients of the deformed cloud: the compiled reference when its build travelled (1e-4 bar), else the CPU
    # oracle (covariance-derived tensors then carry the gcc-vs-nvcc fp32 evaluation-order noise, see ORACLE_GRAD_TOL)
    _, bw = util.run_reference(inp2)
    tol_cov = TOL
    if bw is None:
        _, bw = util.run_oracle(inp2)
        tol_cov = 1e-3
    qt = torch.from_numpy(q).requires_grad_(True)
    torch.nn.functional.normalize(qt, dim=-1).backward(torch.from_numpy(bw["dL_drotations"].astype(np.float64)))
    assert util.rel_l2(tmu.grad.cpu().numpy(), bw["dL_dmeans3D"]) < tol_cov
    assert util.rel_l2(ts.grad.cpu().numpy(), bw["dL_dscales"]) < tol_cov
    assert util.rel_l2(tr.grad.cpu().numpy(), qt.grad.numpy()) < tol_cov


@pytest.mark.parametrize("F", [1, 2, 4, 7, 9, 12, 20, 24, 31])
def test_feature_width_sweep(F):
    """Every run-time feature width dispatches to a padded channel-row layout (NQ in {2,3,5,9}); widths that are not a
    multiple of four take the non-bulk-copy row path.  Ragged image so that partial tiles and partial blocks are hit."""
    inp = util.make_inputs(P=700, W=41, H=27, F=F, seed=100 + F, bg=(0.2, 0.1, 0.7))
    check_stagewise_vs_oracle(inp, 1e-3)


@pytest.mark.parametrize("P,W,H", [(1, 16, 16), (3, 1, 1), (50, 17, 9), (200, 15, 33), (64, 300, 8)])
def test_degenerate_sizes(P, W, H):
    inp = util.make_inputs(P=P, W=W, H=H, F=32, seed=P + W)
    check_stagewise_vs_oracle(inp, 1e-3)


@pytest.mark.parametrize("W,H", [(96, 96), (400, 300), (16, 16)])
def test_tile_launch_order_is_longest_list_first(W, H):
    """tile_order is a permutation of the tiles, sorted by list length descending, ties by tile id (binning.cu
    tile_order_kernel); the blend CTAs take their tile through it, so a wrong entry would leave a tile unrendered -- the
    image parity tests cover that, this one pins the order itself (also for more than 256 tiles... and a single tile)."""
    inp = util.make_inputs(P=4000, W=W, H=H, F=0, seed=17)
    ours, _ = util.run_ours(inp, backward=False)
    lens = (ours["ranges"][:, 1] - ours["ranges"][:, 0]).astype(np.int64)
    T = lens.size
    order = ours["tile_order"].astype(np.int64)
    assert sorted(order.tolist()) == list(range(T))
    expect = sorted(range(T), key=lambda t: (-lens[t], t))
    assert order.tolist() == expect


def test_opaque_scene_early_termination():
    """Large, nearly opaque splats: most pixels stop at T < 1e-4 long before their tile's list ends (forward `done` path,
    backward walks only the first n_contrib records); n_contrib is compared with the oracle inside the stage-wise check."""
    inp = util.make_inputs(P=3000, W=96, H=96, F=32, seed=31, scale0=0.05)
    inp["g"]["opacities"] = np.full_like(inp["g"]["opacities"], 0.95)
    ours, _ = util.run_ours(inp)
    lens = (ours["ranges"][:, 1] - ours["ranges"][:, 0]).astype(np.int64)
    gx = (inp["W"] + 15) // 16
    ys, xs = np.mgrid[0:inp["H"], 0:inp["W"]]
    per_px_len = lens[(ys // 16) * gx + xs // 16].ravel()
    stopped = (ours["final_T"].ravel() < 1e-3) & (ours["n_contrib"].ravel() < per_px_len)
    assert stopped.mean() > 0.3, "the scene was meant to terminate early on a large share of the pixels"
    check_stagewise_vs_oracle(inp, 1e-3)
