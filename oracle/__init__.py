"""CPU oracle for the macroscopic traffic step -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker (or as the
CPU baseline being timed), never as the thing shipped.  The product path
(``mbrl_traffic_b200``) fails loudly when the CUDA library is missing; it never
routes through this package.

Contents
--------
ref_shim.py     imports the UNMODIFIED reference ``lwr.py`` / ``arz.py`` from
                ``/root/reference`` (exists only in the build container).
np_oracle.py    batched NumPy restatement of the reference step, operation
                order preserved (pinned against ref_shim by the golden files).
c_oracle.c      plain-C restatement of the same (threaded CPU baseline).
make_golden.py  regenerates ``tests/golden/*.npz`` from the reference itself.

Parity status: LWR and ARZ are pinned by executing the reference code here
(``tests/golden``).  The non-local model has no reference implementation at
all -> **parity unpinned** for that model.
"""
