"""CPU oracle for the Synthesis-in-Style hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`synthesis_in_style_b200/`) may import, call, link or execute anything in this
directory.  The only permitted users are `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py`, and there only
as the checker / the reported CPU baseline.

The oracle is a torch-CPU (fp32, ATen) restatement of the reference's algorithm;
every function cites the reference file:line it follows
(paths relative to /root/reference, `scf/` = `stylegan_code_finder/`).

Parity pinning: the reference's own tests hold NO golden vectors for this path
(SURVEY.md §4/§8c), so the oracle is pinned against outputs of the reference
itself: `tests/golden/make_golden.py` imports the reference's `model.py`,
`upfirdn2d_native`, `FactorCatalog.pairwise_distance`, `predict_clusters`,
`resize_to_image_size` and `merge_sub_images` from /root/reference in the build
container, asserts the oracle equals them bit-for-bit on seeded inputs, and
commits the resulting vectors under `tests/golden/`.
"""
