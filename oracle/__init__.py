"""CPU oracle for the FeTA spectral hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch/NumPy restatement, on the CPU, of the reference
algorithm (ansonb/FeTA_TMLR) for the path SURVEY.md section 8 names.  Every
function cites the reference file:line it follows.  It is the *checker* for the
CUDA product path in ``feta_tmlr_b200``; nothing in the product may import it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import from here.

PARITY STATUS: pinned by RUNNING THE REFERENCE'S OWN CODE, except for the attention layer.
``tests/golden/make_golden_from_reference.py`` imports the unmodified
``/root/reference/transformer/{ChebNetDynamic,models,data}.py`` under an import shim
(``tests/golden/ref_shim.py``: stand-ins for the PyG-1.7 / torch_scatter / ogb symbols the path touches)
and commits float64 outputs as fixtures; ``tests/test_reference_pin.py`` checks every function of this
package against them (integers bit-exact, floating point <= 1e-6) on all five BASELINE shapes plus the
flag surface.  Still "parity unpinned": ``oracle/layers.py`` -- the source of
``DiffTransformerEncoderLayer`` is missing from the reference tree, so its semantics are
the upstream GraphiT ones cross-checked with the reference's DGL ports -- and the PyG primitives
themselves (torch_geometric==1.7, README.md:24 of the reference), which are restated (twice,
independently: ``oracle/pyg17.py`` and the shim), not executed.  Older substitute pinning stays: (a) an
independent dense-matrix cross-check (``oracle/dense.py``), (b) closed-form known-answer cases, (c)
``torch.autograd.gradcheck`` in fp64.  The integer paths (collate / index builders) are fully determined
by ``transformer/data.py`` and restated exactly.
"""
