"""CPU oracle for the Neural-LAM InteractionNet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``neural_lam_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and there
only as the checker / the timed CPU baseline.

Parity pinning: ``oracle/gen_golden.py`` loads the reference's own
``neural_lam/gnn_layers.py`` + ``neural_lam/utils/networks.py`` UNMODIFIED from
``/root/reference`` (behind ``oracle/pyg_standin.py``, a restatement of the
torch_geometric==2.3.1 ``MessagePassing`` semantics the reference depends on —
PyG itself is absent from the image) and writes the golden vectors under
``tests/golden/``.  ``tests/test_oracle.py`` checks this restatement against
those vectors, and (when ``/root/reference`` is present) against the reference
source directly, including the reference's own ``tests/test_gnn_layers.py``
sections A-H.
"""
