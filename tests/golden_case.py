"""Build (Tree, Patterns-like inputs, Model) from a golden fixture dumped from the reference."""
import numpy as np

from phyml_b200 import model as pmodel
from phyml_b200.tree import Side, Tree

from oracle_backend import load_golden


class GoldenCase:
    def __init__(self, name):
        g = load_golden(name)
        self.g, self.name = g, name
        self.n_otu, self.P, self.ns, self.ncatg = (int(g[k]) for k in ("n_otu", "n_pattern", "ns", "ncatg"))
        self.tree = Tree(self.n_otu, [tuple(x) for x in g["edge_nodes"]], g["edge_l"], list(g["tip_names"]))
        assert (self.tree.left == g["edge_nodes"][:, 0]).all() and (self.tree.rght == g["edge_nodes"][:, 1]).all()
        self.tree.tip_root = int(g["tip_root"])
        assert self.tree.root_edge == int(g["root_edge"])
        self.model = pmodel.from_golden(g, name)
        self.sub = g["sites_sub"]

    def tip_vectors(self, i):
        m = self.g["tip_mask"][i]
        return ((m[:, None] >> np.arange(self.ns, dtype=np.uint32)[None, :]) & 1).astype(np.float64)

    def upload(self, eng, reference_flags=True):
        """Weights, tips (as fp64 vectors, the reference's p_lk_tip_r format) and model."""
        g = self.g
        eng.set_weights(g["wght"], g["invar"])
        for i in range(self.n_otu):
            if reference_flags:
                eng.set_tip_vectors(i, self.tip_vectors(i), g["tip_d_state"][i], g["tip_is_ambigu"][i])
            else:
                eng.set_tip_vectors(i, self.tip_vectors(i))
        eng.set_model(self.model)

    def side(self, handle_or_tip, is_tip):
        return Side(tip=handle_or_tip) if is_tip else Side(clv=handle_or_tip)


ALL_CASES = ["nucleic_hky", "nucleic_gtr_inv", "proteic_lg", "synth_dna_deep", "synth_aa_small", "nucleic_jc_c1"]
