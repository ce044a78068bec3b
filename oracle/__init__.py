"""CPU oracle for the H-Net Mamba ASR encoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``h-net-mamba-asr_b200/dcasr_b200``).  The only legal importers are
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker or the timed CPU arm.

Parity status
-------------
* H-Net stage (``hnet_ref.py``): PINNED.  Checked against the reference's own
  ``dcasr.models.hnet_chunk`` (importable on CPU) by ``tests/golden/make_golden.py``;
  the resulting vectors are committed under ``tests/golden/``.
* Mamba-2 mixer (``mamba2_ref.py``): PARITY UNPINNED BY THE REFERENCE.  The
  arithmetic lives in ``mamba-ssm==2.3.2.post1`` / ``causal-conv1d==1.6.2.post1``
  (requirements.txt:46-47), which are absent from /root/reference and not
  installable here; the reference's tests hold no numeric vectors for it.  The
  restatement follows the published Mamba-2 algorithm and is cross-checked three
  ways (sequential recurrence in fp64, chunked SSD einsum form, and the
  independent ``transformers`` ``Mamba2Mixer.torch_forward``).
"""
