"""GPU parity of the "next" rows of the scope table: fused pre-ops (f2), render()/render_views() (f1),
device-resident cameras (f3).  Everything runs through the C ABI; oracles are the checker only."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ACT_TOL = 2e-6   # row reductions run in a different order than ATen's / numpy's


def _t(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("name", ["f32", "f3_edge", "f0"])
def test_fused_activations_match_torch_golden_and_oracle(name):
    """mgs_activate / mgs_activate_backward vs fixtures made with the reference's PyTorch operators and vs the numpy oracle."""
    import torch
    from manigaussian_b200.gaussian_params import activate_gaussians
    from oracle import activate_oracle as ao
    z = np.load(os.path.join(GOLD, f"activate_{name}.npz"))
    F = int(z["F"])
    leaf = lambda k: _t(z[k]).requires_grad_(True)
    xyz_maps, rot_maps, scale_maps, opacity_maps = leaf("xyz_maps"), leaf("rot_maps"), leaf("scale_maps"), leaf("opacity_maps")
    feature_maps = leaf("feature_maps") if F else None
    means, rot, scales, opac, feat = activate_gaussians(_t(z["xyz"]), rot_maps, scale_maps, opacity_maps, feature_maps, d_means=xyz_maps)
    assert opac.shape == opacity_maps.shape
    fw = ao.forward(z["xyz"], z["rot_maps"], z["scale_maps"], z["opacity_maps"], z["feature_maps"] if F else None, d_means=z["xyz_maps"])
    got = dict(means=means, rot=rot, scales=scales, opac=opac, feature=feat)
    for k in ("means", "rot", "scales", "opac") + (("feature",) if F else ()):
        assert util.rel_l2(got[k].detach().cpu().numpy(), z["out_" + k]) < ACT_TOL, k
        assert util.rel_l2(got[k].detach().cpu().numpy(), fw[k]) < ACT_TOL, k
    loss = sum((got[k] * _t(z["cot_" + k])).sum() for k in ("means", "rot", "scales", "opac") + (("feature",) if F else ()))
    if "next_xyz" in z.files:
        next_xyz, next_rot = leaf("next_xyz"), leaf("next_rot")
        n = activate_gaussians(means.detach(), rot.detach(), scales.detach(), opac.detach(), feat.detach() if F else None,
                               d_means=next_xyz, d_rotations=next_rot, scale_activation=None, opacity_activation=None)
        assert util.rel_l2(n[0].detach().cpu().numpy(), z["next_out_means"]) < ACT_TOL
        assert util.rel_l2(n[1].detach().cpu().numpy(), z["next_out_rot"]) < ACT_TOL
        assert torch.equal(n[2], scales.detach()) and torch.equal(n[3], opac.detach())
        if F:
            # the reference normalises the (already unit) features again on the next-frame render
            assert util.rel_l2(n[4].detach().cpu().numpy(), z["next_out_feature"]) < ACT_TOL
        loss = loss + (n[0] * _t(z["next_cot_means"])).sum() + (n[1] * _t(z["next_cot_rot"])).sum()
    loss.backward()
    pairs = [(xyz_maps, "grad_xyz_maps"), (rot_maps, "grad_rot_maps"), (scale_maps, "grad_scale_maps"), (opacity_maps, "grad_opacity_maps")]
    if F:
        pairs.append((feature_maps, "grad_feature_maps"))
    if "next_xyz" in z.files:
        pairs += [(next_xyz, "grad_next_xyz"), (next_rot, "grad_next_rot")]
    for tns, k in pairs:
        assert tns.grad is not None and tns.grad.shape == tns.shape, k
        assert util.rel_l2(tns.grad.cpu().numpy(), z[k]) < ACT_TOL, k


def test_activation_wide_rows_and_errors():
    """F need not be a power of two or <= 32 for the pre-op kernel; bad shapes raise like the reference's checks."""
    import torch
    from manigaussian_b200.gaussian_params import activate_gaussians
    from manigaussian_b200.gaussian_renderer import normalize_features
    from oracle import activate_oracle as ao
    rng = np.random.default_rng(3)
    for P, F in ((1000, 5), (333, 48), (64, 130), (1, 1)):
        x = rng.normal(size=(P, F)).astype(np.float32)
        g = rng.normal(size=(P, F)).astype(np.float32)
        xt = _t(x).requires_grad_(True)
        y = normalize_features(xt)
        y.backward(_t(g))
        ref = ao.forward(np.zeros((P, 3)), np.ones((P, 4)), np.zeros((P, 3)), np.zeros(P), x)["feature"]
        refg = ao.backward(np.zeros((P, 3)), np.ones((P, 4)), np.zeros((P, 3)), np.zeros(P), x, None, None, None,
                           dict(means=np.zeros((P, 3)), rot=np.zeros((P, 4)), scales=np.zeros((P, 3)), opac=np.zeros(P), feature=g))["feature"]
        assert util.rel_l2(y.detach().cpu().numpy(), ref) < ACT_TOL
        assert util.rel_l2(xt.grad.cpu().numpy(), refg) < ACT_TOL
    with pytest.raises(ValueError):
        activate_gaussians(torch.zeros(4, 3).cuda(), torch.zeros(4, 3).cuda(), torch.zeros(4, 3).cuda(), torch.zeros(4).cuda())
    with pytest.raises(ValueError):
        activate_gaussians(torch.zeros(4, 3).cuda(), torch.zeros(4, 4).cuda(), torch.zeros(4, 3).cuda(), torch.zeros(4).cuda(),
                           scale_activation="softplus")


def _cloud(P, F, seed):
    inp = util.make_inputs(P=P, W=64, H=64, F=F, seed=seed)
    return inp["g"]


def _cams(V, W, H, device="cuda"):
    from manigaussian_b200 import cameras
    sys_path_gold = os.path.join(GOLD, "make_camera_golden.py")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", sys_path_gold)
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    f = W / (2 * np.tan(np.deg2rad(20)))
    intr = np.tile(np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], np.float32), (V, 1, 1))
    extr = np.stack([mk.look_at([0.2 + 1.6 * np.cos(2 * np.pi * v / V), 1.6 * np.sin(2 * np.pi * v / V), 1.5], [0.2, 0, 1.1])
                     for v in range(V)]).astype(np.float32)
    return cameras.build_cameras(intr, extr, W, H, device=device)


def test_render_is_the_reference_pipeline():
    """render() == features / (||f|| + 1e-12) in PyTorch, then GaussianRasterizer -- the reference's render() body
    (gaussian_renderer/__init__.py:17-94) -- bit for bit on the images (the fused normalisation differs by an ulp, so the
    comparison feeds our normalised features to the module path) and within 1e-5 on gradients; and against the CPU oracle."""
    import torch
    from manigaussian_b200 import GaussianRasterizer
    from manigaussian_b200.gaussian_renderer import render, normalize_features
    P, F, W, H = 5000, 32, 96, 80
    g = _cloud(P, F, 21)
    cams = _cams(3, W, H)
    data = {"novel_view": cams.as_novel_view()}
    rng = np.random.default_rng(8)
    raw_feat = (g["feature"] * rng.uniform(0.5, 3.0, (P, 1))).astype(np.float32)   # un-normalised, as the MLP emits them
    leaves = {k: _t(v).requires_grad_(True) for k, v in dict(xyz=g["means3D"], rot=g["rotations"], scale=g["scales"],
                                                             opac=g["opacities"], sh=g["shs"], feat=raw_feat).items()}
    out = render(data, 1, leaves["xyz"], leaves["rot"], leaves["scale"], leaves["opac"], [0, 0, 0], features_color=leaves["sh"],
                 features_language=leaves["feat"], return_depth=True)
    assert set(out) >= {"render", "render_embed", "viewspace_points", "radii"}
    assert out["render"].shape == (3, H, W) and out["render_embed"].shape == (F, H, W) and out["depth"].shape == (H, W)
    ct_c, ct_f = torch.randn(3, H, W, device="cuda"), torch.randn(F, H, W, device="cuda")
    ((out["render"] * ct_c).sum() + (out["render_embed"] * ct_f).sum()).backward()
    # module path with the same normalised features
    nf = normalize_features(leaves["feat"].detach())
    tf = leaves["feat"].detach().clone().requires_grad_(True)
    nf_torch = tf / (tf.norm(dim=-1, keepdim=True) + 1e-12)
    assert util.rel_l2(nf.cpu().numpy(), nf_torch.detach().cpu().numpy()) < ACT_TOL
    s = cams.settings(1, torch.zeros(3, device="cuda"), 1, True)
    l2 = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
    nf_leaf = nf.clone().requires_grad_(True)
    img, emb, radii = GaussianRasterizer(s)(means3D=l2["xyz"], means2D=torch.zeros_like(l2["xyz"]), opacities=l2["opac"], shs=l2["sh"],
                                            language_feature_precomp=nf_leaf, scales=l2["scale"], rotations=l2["rot"])
    assert torch.equal(img, out["render"]) and torch.equal(emb, out["render_embed"]) and torch.equal(radii, out["radii"])
    ((img * ct_c).sum() + (emb * ct_f).sum()).backward()
    nf_torch.backward(nf_leaf.grad)
    for k in ("xyz", "rot", "scale", "opac", "sh"):
        assert util.rel_l2(leaves[k].grad.cpu().numpy(), l2[k].grad.cpu().numpy()) < 1e-5, k
    assert util.rel_l2(leaves["feat"].grad.cpu().numpy(), tf.grad.cpu().numpy()) < 1e-5
    assert out["viewspace_points"].grad is not None and out["viewspace_points"].grad.abs().sum() > 0
    # CPU oracle on the same camera (matrices from the device-resident batch)
    cam = dict(viewmatrix=cams.host["world_view_transform"][1], projmatrix=cams.host["full_proj_transform"][1],
               campos=cams.host["camera_center"][1], tanfovx=cams.tanfovx[1], tanfovy=cams.tanfovy[1])
    g2 = dict(g)
    g2["feature"] = nf.cpu().numpy()
    g2.setdefault("cov3D_precomp", None)
    fw, _ = util.run_oracle(dict(cam=cam, g=g2, ct=None, bg=np.zeros(3, np.float32), P=P, W=W, H=H, F=F), backward=False)
    assert util.rel_l2(out["render"].detach().cpu().numpy(), fw["out_color"]) < 1e-4
    assert util.rel_l2(out["render_embed"].detach().cpu().numpy(), fw["out_feature"]) < 1e-4
    assert np.array_equal(out["radii"].cpu().numpy(), fw["radii"])


def test_render_views_equals_a_loop_of_render():
    """One autograd node for V views: images bit-identical to V calls of render(), gradients equal to the sum of the V
    backward passes (<= 1e-5, atomics' order), per-view screen-space gradients kept apart."""
    import torch
    from manigaussian_b200.gaussian_renderer import render, render_views
    P, F, W, H, V = 8000, 32, 80, 64, 4
    g = _cloud(P, F, 22)
    cams = _cams(V, W, H)
    data = {"novel_view": cams.as_novel_view()}
    mk = lambda: {k: _t(v).requires_grad_(True) for k, v in dict(xyz=g["means3D"], rot=g["rotations"], scale=g["scales"],
                                                                 opac=g["opacities"], sh=g["shs"], feat=g["feature"]).items()}
    a, b = mk(), mk()
    cts = [(torch.randn(3, H, W, device="cuda"), torch.randn(F, H, W, device="cuda"), torch.randn(H, W, device="cuda")) for _ in range(V)]
    outs = render_views(cams, a["xyz"], a["rot"], a["scale"], a["opac"], (0, 0, 0), features_color=a["sh"], features_language=a["feat"],
                        return_depth=True)
    assert outs["render"].shape == (V, 3, H, W) and outs["render_embed"].shape == (V, F, H, W) and outs["radii"].shape == (V, P)
    loss = sum((outs["render"][v] * cts[v][0]).sum() + (outs["render_embed"][v] * cts[v][1]).sum() + (outs["depth"][v] * cts[v][2]).sum()
               for v in range(V))
    loss.backward()
    vsp = []
    for v in range(V):
        o = render(data, v, b["xyz"], b["rot"], b["scale"], b["opac"], [0, 0, 0], features_color=b["sh"], features_language=b["feat"],
                   return_depth=True)
        assert torch.equal(o["render"], outs["render"][v]) and torch.equal(o["render_embed"], outs["render_embed"][v])
        assert torch.equal(o["depth"], outs["depth"][v]) and torch.equal(o["radii"], outs["radii"][v])
        ((o["render"] * cts[v][0]).sum() + (o["render_embed"] * cts[v][1]).sum() + (o["depth"] * cts[v][2]).sum()).backward()
        vsp.append(o["viewspace_points"].grad)
    for k in a:
        assert util.rel_l2(a[k].grad.cpu().numpy(), b[k].grad.cpu().numpy()) < 1e-5, k
    for v in range(V):
        assert util.rel_l2(outs["viewspace_points"].grad[v].cpu().numpy(), vsp[v].cpu().numpy()) < 1e-5
    # a subset of views, precomputed colours, no features
    rgb = torch.rand(P, 3, device="cuda", requires_grad=True)
    sub = render_views(cams, a["xyz"].detach(), a["rot"].detach(), a["scale"].detach(), a["opac"].detach(), (0.1, 0.2, 0.3), pts_rgb=rgb,
                       view_ids=[2, 0])
    assert sub["render"].shape == (2, 3, H, W) and sub["render_embed"] is None
    sub["render"].sum().backward()
    o2 = render(data, 2, a["xyz"].detach(), a["rot"].detach(), a["scale"].detach(), a["opac"].detach(), [0.1, 0.2, 0.3], pts_rgb=rgb.detach())
    assert torch.equal(o2["render"], sub["render"][0])
    assert rgb.grad is not None and rgb.grad.abs().sum() > 0
    # one view: the single-view operator underneath, same layout of the results
    c = mk()
    one = render_views(cams, c["xyz"], c["rot"], c["scale"], c["opac"], (0, 0, 0), features_color=c["sh"], features_language=c["feat"],
                       view_ids=[3], return_depth=True)
    assert one["render"].shape == (1, 3, H, W) and one["depth"].shape == (1, H, W) and one["radii"].shape == (1, P)
    assert torch.equal(one["render"][0], outs["render"][3]) and torch.equal(one["render_embed"][0], outs["render_embed"][3])
    ((one["render"][0] * cts[3][0]).sum() + (one["render_embed"][0] * cts[3][1]).sum() + (one["depth"][0] * cts[3][2]).sum()).backward()
    assert util.rel_l2(one["viewspace_points"].grad[0].cpu().numpy(), vsp[3].cpu().numpy()) < 1e-5


def test_dyna_step_end_to_end_gradients():
    """BASELINE.json configs[3] at test size: raw maps -> fused activations -> current-frame views; detached + deformation
    offsets -> next-frame views; one backward.  Gradients w.r.t. every raw map and both offsets vs the same graph built from
    the reference's PyTorch operators around the single-view rasterizer module."""
    import torch
    from manigaussian_b200 import GaussianRasterizer
    from manigaussian_b200.gaussian_params import activate_gaussians
    from manigaussian_b200.gaussian_renderer import render_views
    P, F, W, H, V = 6000, 32, 64, 64, 2
    g = _cloud(P, F, 23)
    cams = _cams(V, W, H)
    rng = np.random.default_rng(4)
    raw = dict(xyz=g["means3D"], xyz_maps=rng.normal(0, 0.01, (P, 3)), rot_maps=g["rotations"] * rng.uniform(0.5, 2, (P, 1)),
               scale_maps=np.log(g["scales"]) + rng.normal(0, 0.05, (P, 3)), opacity_maps=np.log(g["opacities"] / (1 - g["opacities"])),
               sh=g["shs"], feature_maps=g["feature"] * rng.uniform(0.5, 2, (P, 1)), next_xyz=rng.normal(0, 0.01, (P, 3)),
               next_rot=rng.normal(0, 0.05, (P, 4)))
    mk = lambda: {k: _t(np.asarray(v, np.float32)).requires_grad_(k != "xyz") for k, v in raw.items()}
    cts = [[(torch.randn(3, H, W, device="cuda"), torch.randn(F, H, W, device="cuda")) for _ in range(V)] for _ in range(2)]

    a = mk()
    cur = activate_gaussians(a["xyz"], a["rot_maps"], a["scale_maps"], a["opacity_maps"], a["feature_maps"], d_means=a["xyz_maps"],
                             normalize_feature=False)
    nxt = activate_gaussians(cur[0].detach(), cur[1].detach(), cur[2].detach(), cur[3].detach(), None, d_means=a["next_xyz"],
                             d_rotations=a["next_rot"], scale_activation=None, opacity_activation=None)
    o_cur = render_views(cams, cur[0], cur[1], cur[2], cur[3], features_color=a["sh"], features_language=cur[4])
    o_nxt = render_views(cams, nxt[0], nxt[1], nxt[2], nxt[3], features_color=a["sh"].detach(), features_language=cur[4].detach())
    loss = sum((o["render"][v] * c[v][0]).sum() + (o["render_embed"][v] * c[v][1]).sum() for o, c in ((o_cur, cts[0]), (o_nxt, cts[1]))
               for v in range(V))
    loss.backward()

    b = mk()
    scales = torch.clamp_max(torch.exp(b["scale_maps"]), 0.05)
    means = b["xyz"] + b["xyz_maps"]
    rots = torch.nn.functional.normalize(b["rot_maps"], dim=-1)
    opac = torch.sigmoid(b["opacity_maps"])
    n_means = means.detach() + b["next_xyz"]
    n_rots = torch.nn.functional.normalize(rots.detach() + b["next_rot"], dim=-1)

    def ref_render(v, m, r, s, o, sh, f):
        f = f / (f.norm(dim=-1, keepdim=True) + 1e-12)
        st = cams.settings(v, torch.zeros(3, device="cuda"), 1, True)
        return GaussianRasterizer(st)(means3D=m, means2D=torch.zeros_like(m), opacities=o, shs=sh, language_feature_precomp=f, scales=s, rotations=r)
    loss_b = 0
    for v in range(V):
        img, emb, _ = ref_render(v, means, rots, scales, opac, b["sh"], b["feature_maps"])
        loss_b = loss_b + (img * cts[0][v][0]).sum() + (emb * cts[0][v][1]).sum()
        img, emb, _ = ref_render(v, n_means, n_rots, scales.detach(), opac.detach(), b["sh"].detach(), b["feature_maps"].detach())
        loss_b = loss_b + (img * cts[1][v][0]).sum() + (emb * cts[1][v][1]).sum()
    loss_b.backward()
    assert abs(float(loss) - float(loss_b)) <= 1e-4 * abs(float(loss_b)) + 1e-3
    for k in ("xyz_maps", "rot_maps", "scale_maps", "opacity_maps", "sh", "feature_maps", "next_xyz", "next_rot"):
        assert a[k].grad is not None, k
        assert util.rel_l2(a[k].grad.cpu().numpy(), b[k].grad.cpu().numpy()) < 2e-5, k


def test_step_is_cuda_graph_capturable():
    """The multi-view step never synchronises with the host and allocates only through torch's allocator, so forward +
    loss + backward can be captured ONCE with torch.cuda.graph and replayed (the reference's forward blocks on a device->host
    copy of its instance count, rasterizer_impl.cu:284, and cannot).  Replays must reproduce the eager gradients, also after
    the inputs changed in place."""
    import torch
    from manigaussian_b200.gaussian_renderer import render_views
    P, F, W, H, V = 12000, 32, 80, 64, 3
    g = _cloud(P, F, 29)
    cams = _cams(V, W, H)
    names = ("means3D", "rotations", "scales", "opacities", "shs", "feature")
    L = {k: _t(np.asarray(g[k], np.float32)).requires_grad_(True) for k in names}
    cts = (torch.randn(V, 3, H, W, device="cuda"), torch.randn(V, F, H, W, device="cuda"))

    def fwd_bwd():
        o = render_views(cams, L["means3D"], L["rotations"], L["scales"], L["opacities"], features_color=L["shs"],
                         features_language=L["feature"])
        loss = (o["render"] * cts[0]).sum() + (o["render_embed"] * cts[1]).sum()
        loss.backward()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            for v in L.values():
                v.grad = None
            fwd_bwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    eager = {k: v.grad.clone() for k, v in L.items()}
    for v in L.values():
        v.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = fwd_bwd()
    for rep in range(2):
        graph.replay()
    torch.cuda.synchronize()
    for k in names:
        assert util.rel_l2(L[k].grad.cpu().numpy(), eager[k].cpu().numpy()) < 1e-5, k
    # new input values in the same buffers: the replay must follow them
    with torch.no_grad():
        L["means3D"].add_(0.01)
    graph.replay()
    torch.cuda.synchronize()
    moved = {k: v.grad.clone() for k, v in L.items()}
    for v in L.values():
        v.grad = None
    fwd_bwd()
    torch.cuda.synchronize()
    for k in names:
        assert util.rel_l2(moved[k].cpu().numpy(), L[k].grad.cpu().numpy()) < 1e-5, k


@pytest.mark.parametrize("name", ["f32", "f3"])
def test_loss_heads_kernel_matches_reference_golden(name):
    """mgs_loss_heads (the device code the forward blend's epilogue shares, loss_heads.cuh) on the golden images produced
    with the REFERENCE's own l2_loss / cosine_loss (tests/golden/make_loss_golden.py): loss values and both cotangent planes."""
    import torch
    from manigaussian_b200 import _binding as b
    z = np.load(os.path.join(GOLD, "loss_heads.npz"))
    render, gt, embed, gt_e = (_t(z[f"{name}_{k}"])[None].contiguous() for k in ("render", "gt", "embed", "gt_embed"))
    F, H, W = embed.shape[1:]
    cot_c, cot_f = torch.empty_like(render), torch.empty_like(embed)
    acc = torch.empty((1, 2), device="cuda")
    b.check(b.lib().mgs_loss_heads(1, F, H * W, render.data_ptr(), embed.data_ptr(), gt.data_ptr(), gt_e.data_ptr(), cot_c.data_ptr(),
                                   cot_f.data_ptr(), acc.data_ptr(), torch.cuda.current_stream().cuda_stream), "mgs_loss_heads")
    torch.cuda.synchronize()
    loss_rgb = float(acc[0, 0]) / (3 * H * W)
    loss_embed = 1.0 - float(acc[0, 1]) / (H * W)
    assert abs(loss_rgb - float(z[f"{name}_loss_rgb"])) <= 2e-6 * abs(float(z[f"{name}_loss_rgb"]))
    assert abs(loss_embed - float(z[f"{name}_loss_embed"])) <= 2e-6
    assert util.rel_l2(cot_c[0].cpu().numpy(), z[f"{name}_d_render"]) < 2e-6
    assert util.rel_l2(cot_f[0].cpu().numpy(), z[f"{name}_d_embed"]) < 1e-5


@pytest.mark.parametrize("F,lam", [(32, 0.01), (3, 1.0), (0, 0.0)])
def test_loss_heads_fused_into_the_blend_epilogue(F, lam):
    """render_views(targets=...) -- the L2 colour head and the cosine embedding head evaluated in the forward blend's
    epilogue, their cotangents consumed by the backward blend -- against the same objective built from the reference's
    PyTorch expressions (loss.py:12-23 as combined in neural_rendering.py:300-318) on the unfused render."""
    import torch
    from manigaussian_b200.gaussian_renderer import render_views
    P, W, H, V = 9000, 72, 56, 3
    g = _cloud(P, max(F, 1), 37)
    cams = _cams(V, W, H)
    names = ("means3D", "rotations", "scales", "opacities", "shs") + (("feature",) if F else ())
    tgt_rgb = torch.rand(V, 3, H, W, device="cuda")
    tgt_emb = torch.randn(V, F, H, W, device="cuda") if F else None

    def leaves():
        d = {k: _t(np.asarray(g[k], np.float32)).requires_grad_(True) for k in names}
        if F and F != g["feature"].shape[1]:
            d["feature"] = _t(np.asarray(g["feature"][:, :F], np.float32)).requires_grad_(True)
        return d

    a = leaves()
    o = render_views(cams, a["means3D"], a["rotations"], a["scales"], a["opacities"], features_color=a["shs"],
                     features_language=a.get("feature"), targets={"rgb": tgt_rgb, "embed": tgt_emb})
    loss_a = o["loss_rgb"].sum() + lam * o["loss_embed"].sum()
    loss_a.backward()

    b = leaves()
    r = render_views(cams, b["means3D"], b["rotations"], b["scales"], b["opacities"], features_color=b["shs"],
                     features_language=b.get("feature"))
    loss_b = 0
    for v in range(V):
        loss_b = loss_b + ((r["render"][v] - tgt_rgb[v]) ** 2).mean()
        if F:
            cs = torch.nn.functional.cosine_similarity(r["render_embed"][v].permute(1, 2, 0), tgt_emb[v].permute(1, 2, 0), dim=-1)
            loss_b = loss_b + lam * (1 - cs.mean())
    loss_b.backward()
    assert torch.equal(o["render"], r["render"].detach())
    assert abs(float(loss_a) - float(loss_b)) <= 1e-5 * abs(float(loss_b))
    for k in names:
        assert a[k].grad is not None, k
        assert util.rel_l2(a[k].grad.cpu().numpy(), b[k].grad.cpu().numpy()) < 2e-5, k
    assert util.rel_l2(o["viewspace_points"].grad.cpu().numpy(), r["viewspace_points"].grad.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("path", ["module", "render_views", "render_views_depth", "loss_heads"])
def test_steps_do_not_accumulate_device_memory(path):
    """A training loop must come back to the same allocated bytes after every step.  An autograd Function that keeps one of
    its OUTPUTS on ctx builds the cycle output -> grad_fn -> ctx -> output, which neither reference counting nor the
    garbage collector frees: every step's images, state buffers and -- through the graph -- its input leaves stay alive."""
    import gc
    import torch
    from manigaussian_b200 import GaussianRasterizer
    from manigaussian_b200.gaussian_renderer import render_views
    P, F, W, H, V = 6000, 32, 64, 48, 2
    g = _cloud(P, F, 33)
    cams = _cams(V, W, H)
    host = {k: np.asarray(g[k], np.float32) for k in ("means3D", "rotations", "scales", "opacities", "shs", "feature")}
    tgt = (torch.rand(V, 3, H, W, device="cuda"), torch.randn(V, F, H, W, device="cuda"))

    def step():
        L = {k: _t(v).requires_grad_(True) for k, v in host.items()}  # fresh leaves every step, like a network's outputs
        if path == "module":
            s = cams.settings(0, torch.zeros(3, device="cuda"), 1, True)
            img, emb, _ = GaussianRasterizer(s)(means3D=L["means3D"], means2D=torch.zeros_like(L["means3D"]), opacities=L["opacities"],
                                                shs=L["shs"], language_feature_precomp=L["feature"], scales=L["scales"],
                                                rotations=L["rotations"])
            loss = (img * tgt[0][0]).sum() + (emb * tgt[1][0]).sum()
        elif path == "loss_heads":
            o = render_views(cams, L["means3D"], L["rotations"], L["scales"], L["opacities"], features_color=L["shs"],
                             features_language=L["feature"], targets={"rgb": tgt[0], "embed": tgt[1]})
            loss = o["loss_rgb"].sum() + 0.01 * o["loss_embed"].sum()
        else:
            o = render_views(cams, L["means3D"], L["rotations"], L["scales"], L["opacities"], features_color=L["shs"],
                             features_language=L["feature"], return_depth=(path == "render_views_depth"))
            loss = (o["render"] * tgt[0]).sum() + (o["render_embed"] * tgt[1]).sum()
            if path == "render_views_depth":
                loss = loss + o["depth"].sum()
        loss.backward()
        return float(loss)

    gc.disable()
    try:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        grown = torch.cuda.memory_allocated() - base
    finally:
        gc.enable()
    assert grown == 0, f"{grown} bytes stayed allocated after 4 more steps"
