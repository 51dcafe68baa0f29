"""Oracle (oracle/postprocess.py) vs golden vectors produced by the reference's OWN function texts (utils.py:48-59, 403-776;
sample.py:316-329), exec()'d verbatim by tests/golden/make_golden_post.py on synthetic closed B-reps: vertex / edge ids and
adjacency exact (same numbering), coordinates to fp32 round-off."""
import os

import numpy as np
import pytest

from make_golden_post import CASES, Z_THRESHOLD, select_cad, synth_cad
from oracle import postprocess as OP

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_golden.npz"))


def unpack(off, flat):
    return [list(flat[off[i]:off[i + 1]]) for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_topology_matches_reference(name):
    c = select_cad(synth_cad(*CASES[name]))
    ends = OP.edge_endpoints(c["edge_pos_cad"], c["edge_ncs_cad"], c["edge_mask_cad"])
    assert np.array_equal(np.concatenate(ends).astype(np.float32), GOLD[f"{name}|edgeV_bbox"])
    uv, vd = OP.detect_shared_vertex(c["edgeV_cad"], c["edge_mask_cad"], ends)
    assert [sorted(vd[k]) for k in range(len(vd))] == [sorted(x) for x in unpack(GOLD[f"{name}|vertex_dict_off"],
                                                                                 GOLD[f"{name}|vertex_dict"])]
    assert np.allclose(uv, GOLD[f"{name}|unique_vertices"], atol=1e-6)
    uf, ue, fea, eva = OP.detect_shared_edge(uv, vd, c["edge_z_cad"], c["surf_z_cad"], Z_THRESHOLD, c["edge_mask_cad"])
    assert np.array_equal(eva, GOLD[f"{name}|EdgeVertexAdj"])
    assert [list(r) for r in fea] == unpack(GOLD[f"{name}|FaceEdgeAdj_off"], GOLD[f"{name}|FaceEdgeAdj"])
    assert np.array_equal(ue, GOLD[f"{name}|unique_edges"])


def test_oracle_joint_optimize_matches_reference():
    name = "box_a"
    c = select_cad(synth_cad(*CASES[name]))
    uv = GOLD[f"{name}|unique_vertices"]
    eva = GOLD[f"{name}|EdgeVertexAdj"]
    fea = unpack(GOLD[f"{name}|FaceEdgeAdj_off"], GOLD[f"{name}|FaceEdgeAdj"])
    edge_ncs_u = c["edge_ncs_cad"][~c["edge_mask_cad"]][GOLD[f"{name}|unique_edge_ids"]]
    surf_wcs, edge_wcs = OP.joint_optimize(c["surf_ncs_cad"], edge_ncs_u, c["surf_pos_cad"], uv, eva, fea, len(edge_ncs_u),
                                           len(c["surf_ncs_cad"]))
    assert np.abs(edge_wcs - GOLD[f"{name}|edge_wcs"]).max() < 1e-6
    assert np.abs(surf_wcs - GOLD[f"{name}|surf_wcs"]).max() < 1e-5
