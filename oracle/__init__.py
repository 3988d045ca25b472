"""CPU oracle for the superpixel-segmented scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mulactseg_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU baseline -- never as the product path.

What it is: a plain torch-CPU / numpy restatement of the reference's algorithm
for the path named by BASELINE.json ``north_star`` (acquisition selectors,
multi-hot / MIL losses, prototype pseudo-labeller, region selection).  Each
function cites the reference file:line it follows (paths relative to the
reference checkout, sehyun03/MulActSeg).

Third-party arithmetic that is NOT in the reference tree: ``torch_scatter``
(conda ``pytorch-scatter=2.0.9=py38_torch_1.11.0_cu113``, ``actsegmul.yml:99``).
``oracle/scatter_ref.py`` restates its published semantics.  The reference has
no tests / golden vectors for that boundary, so parity is pinned differently:
``oracle/gen_golden.py`` imports the UNMODIFIED reference classes from
``/root/reference`` (in the build container, where that checkout exists), runs
them over ``scatter_ref`` on seeded synthetic tensors and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors on any box.
"""
