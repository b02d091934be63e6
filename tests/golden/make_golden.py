"""Generate golden vectors from the REFERENCE's own independent oracle.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``/root/reference/fvdb/utils/tests/convolution_semantics_oracle.py`` by file path (that
module imports only torch; ``import fvdb`` itself fails here because ``_fvdb_cpp`` cannot be built)
and records, for a matrix of geometries and signed coordinate sets taken from the reference tests
(tests/unit/test_conv_semantics.py, tests/unit/test_conv_semantics_integration.py:129-148,170-242):

* ``relation_edges`` / ``forward_degrees`` / ``forward_support`` / ``transpose_support``;
* ``dense_forward_oracle`` / ``dense_transpose_oracle`` values at every target coordinate plus the
  gradients w.r.t. features and weights for a fixed probe (fp64, seeds as in the reference test).

Outputs: tests/golden/semantics_golden.json.gz and tests/golden/dense_golden.npz (committed).
"""

import gzip
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import torch

REFERENCE_ORACLE = Path("/root/reference/fvdb/utils/tests/convolution_semantics_oracle.py")
HERE = Path(__file__).resolve().parent


def _load_reference_oracle():
    spec = importlib.util.spec_from_file_location("reference_convolution_semantics_oracle", REFERENCE_ORACLE)
    module = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = module
    spec.loader.exec_module(module)
    return module


UNIFORM = [((k, k, k), (s, s, s)) for k in range(1, 7) for s in range(1, 6)]
MIXED = [((2, 3, 4), (1, 2, 3)), ((5, 2, 3), (4, 2, 1)), ((3, 4, 2), (2, 3, 4)), ((4, 1, 1), (4, 1, 1)), ((4, 3, 2), (3, 2, 1)), ((3, 3, 3), (4, 4, 4))]
COORD_SETS = {
    "signed5_a": [(-4, -1, 0), (-1, 0, 1), (0, 2, -3), (3, -2, 4), (5, 1, -1)],
    "signed5_b": [(-3, 0, 0), (-1, 1, 0), (0, -1, 1), (2, 0, -1), (5, 2, 1)],
    "signed5_c": [(-5, -1, 0), (-1, 0, 1), (0, 2, -3), (3, -2, 4), (8, 1, -1)],
    "line16": [(c, 0, 0) for c in range(16)],
}
DENSE_GEOMETRIES = [((4, 2, 2), (4, 2, 2)), ((2, 3, 4), (1, 2, 3)), ((2, 1, 3), (3, 2, 4)), ((3, 3, 3), (1, 1, 1)), ((3, 3, 3), (2, 2, 2)), ((2, 2, 2), (2, 2, 2))]


def main() -> None:
    ref = _load_reference_oracle()
    semantics = []
    for kernel, stride in UNIFORM + MIXED:
        relation = ref.ConvolutionRelation(kernel, stride)
        for name, coords in COORD_SETS.items():
            if name == "line16" and (kernel, stride) != ((4, 1, 1), (4, 1, 1)):
                continue
            edges = ref.relation_edges(coords, relation)
            support = sorted(ref.forward_support(coords, relation))
            semantics.append(
                {
                    "kernel_size": list(kernel),
                    "stride": list(stride),
                    "coords_name": name,
                    "fine": [list(c) for c in coords],
                    "p_before": list(relation.p_before),
                    "p_after": list(relation.p_after),
                    "offsets_axis0": [relation.offset((t, 0, 0))[0] for t in range(kernel[0])],
                    "edges": [[list(e.fine), list(e.coarse), list(e.tap)] for e in edges],
                    "forward_degrees": [[list(c), d] for c, d in sorted(ref.forward_degrees(coords, relation).items())],
                    "forward_support": [list(c) for c in support],
                    "transpose_support_of_forward": [list(c) for c in sorted(ref.transpose_support(support, relation))],
                }
            )
    with gzip.open(HERE / "semantics_golden.json.gz", "wt", compresslevel=9) as f:
        json.dump({"source": str(REFERENCE_ORACLE), "cases": semantics}, f, separators=(",", ":"))

    dense = {}
    meta = []
    for case_id, (kernel, stride) in enumerate(DENSE_GEOMETRIES):
        relation = ref.ConvolutionRelation(kernel, stride)
        # coordinate recipe of tests/unit/test_conv_semantics_integration.py:186-189: every tap is exercised
        coords = sorted({relation.fine_from_coarse(coarse, tap) for coarse in ((0, 0, 0), (-2, 1, -1)) for tap in relation.taps()})
        for transposed in (False, True):
            generator = torch.Generator().manual_seed(668 + int(transposed))
            features = torch.randn((len(coords), 2), generator=generator, dtype=torch.float64).requires_grad_()
            count = 3 * 2 * kernel[0] * kernel[1] * kernel[2]
            weights = (torch.arange(1, count + 1, dtype=torch.float64).reshape(3, 2, *kernel) / count).requires_grad_()
            if transposed:
                result = ref.dense_transpose_oracle(coords, features, weights, relation)
                targets = sorted(ref.transpose_support(coords, relation))
            else:
                result = ref.dense_forward_oracle(coords, features, weights, relation)
                targets = sorted(ref.forward_support(coords, relation))
            values = torch.stack([result.value_at(t) for t in targets])
            probe = torch.arange(1, values.numel() + 1, dtype=torch.float64).reshape_as(values)
            grad_f, grad_w = torch.autograd.grad(torch.sum(values * probe), (features, weights))
            key = f"c{case_id}_{'t' if transposed else 'f'}"
            dense[key + "_source"] = np.asarray(coords, dtype=np.int32)
            dense[key + "_target"] = np.asarray(targets, dtype=np.int32)
            dense[key + "_features"] = features.detach().numpy()
            dense[key + "_weights"] = weights.detach().numpy()
            dense[key + "_values"] = values.detach().numpy()
            dense[key + "_grad_features"] = grad_f.numpy()
            dense[key + "_grad_weights"] = grad_w.numpy()
            meta.append({"key": key, "kernel_size": list(kernel), "stride": list(stride), "transposed": transposed})
    dense["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(HERE / "dense_golden.npz", **dense)
    print(f"wrote {len(semantics)} semantics cases, {len(meta)} dense cases")


if __name__ == "__main__":
    main()
