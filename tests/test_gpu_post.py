"""GPU: the post-decode geometry glue (brepgen_b200/postprocess.py over csrc/geom.cu) against golden vectors produced by the
reference's OWN function texts (tests/golden/make_golden_post.py: utils.py:48-59, 403-776 and sample.py:316-329 exec()'d
verbatim on synthetic closed B-reps).  Topology (vertex groups, EdgeVertexAdj, FaceEdgeAdj, unique edges) must be IDENTICAL
including the numbering; coordinates to fp32 round-off; the 200-step AdamW surface fit to 2e-4 (its update has magnitude
~lr = 1e-3 per step whatever the gradient's size, so summation-order noise is amplified near convergence)."""
import os

import numpy as np
import pytest
import torch

from make_golden_post import CASES, Z_THRESHOLD, select_cad, synth_cad
from oracle import postprocess as OP

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_golden.npz"))


def unpack(off, flat):
    return [list(flat[off[i]:off[i + 1]]) for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", list(CASES))
def test_topology_and_edges_match_reference(name):
    from brepgen_b200 import postprocess as P
    c = select_cad(synth_cad(*CASES[name]))
    ends = P.edge_endpoints(c["edge_pos_cad"], c["edge_ncs_cad"], c["edge_mask_cad"])
    assert [e.shape for e in ends] == [(4, 2, 3)] * 6
    assert np.abs(np.concatenate(ends) - GOLD[f"{name}|edgeV_bbox"]).max() <= 1e-7
    uv, vd = P.detect_shared_vertex(c["edgeV_cad"], c["edge_mask_cad"], ends)
    want = unpack(GOLD[f"{name}|vertex_dict_off"], GOLD[f"{name}|vertex_dict"])
    assert [sorted(int(x) for x in vd[k]) for k in range(len(vd))] == [sorted(int(x) for x in w) for w in want]
    assert np.abs(uv - GOLD[f"{name}|unique_vertices"]).max() < 1e-6
    uf, ue, fea, eva = P.detect_shared_edge(uv, vd, c["edge_z_cad"], c["surf_z_cad"], Z_THRESHOLD, c["edge_mask_cad"])
    assert np.array_equal(eva, GOLD[f"{name}|EdgeVertexAdj"])
    assert [[int(x) for x in r] for r in fea] == [[int(x) for x in r] for r in unpack(GOLD[f"{name}|FaceEdgeAdj_off"],
                                                                                        GOLD[f"{name}|FaceEdgeAdj"])]
    assert np.array_equal(ue, GOLD[f"{name}|unique_edges"])
    # single-face entry point (utils.py:403-421) against the oracle
    for f in range(6):
        assert np.array_equal(P.edge2loop(ends[f]), OP.edge2loop(ends[f]))


@pytest.mark.parametrize("name", list(CASES))
def test_joint_optimize_matches_reference(name):
    from brepgen_b200 import postprocess as P
    c = select_cad(synth_cad(*CASES[name]))
    uv, eva = GOLD[f"{name}|unique_vertices"], GOLD[f"{name}|EdgeVertexAdj"]
    fea = unpack(GOLD[f"{name}|FaceEdgeAdj_off"], GOLD[f"{name}|FaceEdgeAdj"])
    edge_ncs_u = c["edge_ncs_cad"][~c["edge_mask_cad"]][GOLD[f"{name}|unique_edge_ids"]]
    surf_wcs, edge_wcs = P.joint_optimize(c["surf_ncs_cad"], edge_ncs_u, c["surf_pos_cad"], uv, eva, fea, len(edge_ncs_u),
                                          len(c["surf_ncs_cad"]))
    assert surf_wcs.shape == GOLD[f"{name}|surf_wcs"].shape and edge_wcs.shape == GOLD[f"{name}|edge_wcs"].shape
    e_err = np.abs(edge_wcs - GOLD[f"{name}|edge_wcs"]).max()
    s_err = np.abs(surf_wcs - GOLD[f"{name}|surf_wcs"]).max()
    # zero iterations = the initial surfaces: pins bg_surf_init separately from the optimiser
    s0, _ = P.joint_optimize(c["surf_ncs_cad"], edge_ncs_u, c["surf_pos_cad"], uv, eva, fea, len(edge_ncs_u),
                             len(c["surf_ncs_cad"]), iters=1)
    init = OP.init_surfaces(c["surf_ncs_cad"], c["surf_pos_cad"], GOLD[f"{name}|edge_wcs"], fea)
    i_err = np.abs(s0 - init).max()
    print(f"joint_optimize {name}: edge_wcs max abs err {e_err:.2e}, surf init {i_err:.2e}, surf after 200 AdamW steps {s_err:.2e}")
    assert e_err < 2e-6 and i_err < 2e-6 and s_err < 2e-4


def test_surface_fit_batched_over_many_faces():
    """one launch, 6 x 40 faces of 40 CADs: every face gets the same result as when its CAD is fitted alone"""
    from brepgen_b200 import postprocess as P
    name = "box_b"
    c = select_cad(synth_cad(*CASES[name]))
    uv, eva = GOLD[f"{name}|unique_vertices"], GOLD[f"{name}|EdgeVertexAdj"]
    fea = unpack(GOLD[f"{name}|FaceEdgeAdj_off"], GOLD[f"{name}|FaceEdgeAdj"])
    edge_ncs_u = c["edge_ncs_cad"][~c["edge_mask_cad"]][GOLD[f"{name}|unique_edge_ids"]]
    one, _ = P.joint_optimize(c["surf_ncs_cad"], edge_ncs_u, c["surf_pos_cad"], uv, eva, fea, 12, 6)
    # replicate the CAD 40 times inside ONE call: adjacency offsets shifted, per-face 1/6 weights via a 6-face mean
    reps = 40
    many_fea = [[e + 12 * r for e in row] for r in range(reps) for row in fea]
    from brepgen_b200 import _ffi
    dev = torch.device("cuda")
    t = lambda a, dt=np.float32: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev)
    e_ncs = t(np.tile(edge_ncs_u, (reps, 1, 1)))
    vse = t(np.tile(uv[eva], (reps, 1, 1)))
    ewcs = torch.empty(12 * reps, 32, 3, device=dev)
    _ffi.check(_ffi.lib().bg_edge_fit(e_ncs.data_ptr(), vse.data_ptr(), 12 * reps, ewcs.data_ptr(), _ffi.current_stream()))
    off = t(np.concatenate([[0], np.cumsum([len(a) for a in many_fea])]), np.int32)
    adj = t(np.concatenate(many_fea), np.int32)
    init = torch.empty(6 * reps, 1024, 3, device=dev)
    s_ncs = t(np.tile(c["surf_ncs_cad"].reshape(6, -1, 3), (reps, 1, 1)))
    s_pos = t(np.tile(c["surf_pos_cad"], (reps, 1)))
    _ffi.check(_ffi.lib().bg_surf_init(s_ncs.data_ptr(), s_pos.data_ptr(), ewcs.data_ptr(), off.data_ptr(), adj.data_ptr(), 6 * reps,
                                      init.data_ptr(), _ffi.current_stream()))
    inv = torch.full((6 * reps,), 1.0 / 6, device=dev)
    out = torch.empty_like(init)
    _ffi.check(_ffi.lib().bg_surf_offset_opt(init.data_ptr(), ewcs.data_ptr(), off.data_ptr(), adj.data_ptr(), inv.data_ptr(),
                                            6 * reps, 4, 200, 1e-3, 0.95, 0.999, 1e-8, 1e-6, out.data_ptr(), None,
                                            _ffi.current_stream()))
    got = out.reshape(reps, 6, 32, 32, 3).cpu().numpy()
    assert all(np.array_equal(got[r], one) for r in range(reps))


def test_postprocess_cad_end_to_end():
    """the whole loop body sample.py:303-355 (up to construct_brep) for one CAD: product pipeline (GPU cores + drop-in decoders)
    against the oracle pipeline (numpy topology + oracle VAE decode + torch-CPU joint_optimize)"""
    from brepgen_b200 import postprocess as P
    from brepgen_b200.spec import edge_decoder_spec, surf_decoder_spec
    from brepgen_b200.synth import synth_state_dict
    from brepgen_b200.vae import build_synthetic_decoders
    from oracle import vae as V
    d = synth_cad(*CASES["box_c"])
    sv, ev = build_synthetic_decoders(torch.device("cuda"), seed=2)
    surf_wcs, edge_wcs, fea, eva, uv = P.postprocess_cad(sv, ev, d["surfPos"], d["surfZ"], d["surfMask"], d["edge_pos"], d["edge_ncs"],
                                                         d["edgeV"], d["edge_z"], d["edge_mask"], Z_THRESHOLD)
    c = select_cad(d)
    ends = OP.edge_endpoints(c["edge_pos_cad"], c["edge_ncs_cad"], c["edge_mask_cad"])
    ouv, ovd = OP.detect_shared_vertex(c["edgeV_cad"], c["edge_mask_cad"], ends)
    ouf, oue, ofea, oeva = OP.detect_shared_edge(ouv, ovd, c["edge_z_cad"], c["surf_z_cad"], Z_THRESHOLD, c["edge_mask_cad"])
    assert np.array_equal(eva, oeva) and [list(map(int, r)) for r in fea] == [list(map(int, r)) for r in ofea]
    sd_s, sd_e = synth_state_dict(surf_decoder_spec(), 2), synth_state_dict(edge_decoder_spec(), 2)
    with torch.no_grad():
        zf = torch.as_tensor(ouf).unflatten(-1, (16, 3)).permute(0, 2, 1).unflatten(-1, (4, 4))
        o_surf = V.surf_decode(sd_s, zf).permute(0, 2, 3, 1).numpy()
        ze = torch.as_tensor(oue).unflatten(-1, (4, 3)).permute(0, 2, 1)
        o_edge = V.edge_decode(sd_e, ze).permute(0, 2, 1).numpy()
    o_swcs, o_ewcs = OP.joint_optimize(o_surf, o_edge, c["surf_pos_cad"], ouv, oeva, ofea, len(o_edge), len(o_surf))
    e_err = np.abs(edge_wcs - o_ewcs).max() / np.abs(o_ewcs).max()
    s_err = np.abs(surf_wcs - o_swcs).max() / np.abs(o_swcs).max()
    print(f"postprocess_cad: edge_wcs rel max err {e_err:.2e}, surf_wcs {s_err:.2e}")
    assert surf_wcs.shape == (6, 32, 32, 3) and edge_wcs.shape == (12, 32, 3)
    assert e_err < 1e-3 and s_err < 2e-3
