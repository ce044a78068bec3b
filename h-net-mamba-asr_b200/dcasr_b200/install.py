"""Drop-in installation under the reference package.

    import dcasr_b200; dcasr_b200.install()      # before `import dcasr.tasks...`

After this, ``from mamba_ssm import Mamba2`` (reference src/dcasr/models/mamba_block.py:12) resolves to
the B200 mixer, and ``dcasr.models.hnet_chunk`` / ``.mamba_block`` / ``.encoder`` expose the B200 classes,
so ``dcasr.tasks.asr_task.build_model`` (reference :129-146), ``scripts/train.py`` and
``scripts/decode.py`` construct and run the CUDA hot path with no source change.  ``Trainer._any_rank_oom`` (reference
training/trainer.py:200-208) is rebound to a host-side collective (trainer_sync.py): same flag on every rank, no
``.item()`` on the CUDA stream after every micro-batch.
"""
from __future__ import annotations

import sys
import types


def install(patch_dcasr: bool = True, patch_trainer: bool = True) -> None:
    from . import encoder, hnet_chunk, mamba_block

    shim = types.ModuleType("mamba_ssm")
    shim.Mamba2 = mamba_block.Mamba2
    shim.__doc__ = "hnet_b200 stand-in for mamba_ssm (Mamba2 only)"
    sys.modules["mamba_ssm"] = shim
    if not patch_dcasr:
        return
    try:
        import dcasr.models as ref_models                      # the reference package, if importable
    except Exception:
        return
    import dcasr.models.hnet_chunk as ref_chunk
    for name in ("ChunkOutput", "RoutingModule", "DynamicChunker", "ratio_loss"):
        setattr(ref_chunk, name, getattr(hnet_chunk, name))
        setattr(ref_models, name, getattr(hnet_chunk, name))
    try:                                                        # `chunker: fixed` resolves to the CUDA FixedPoolChunker too
        import dcasr.models.fixed_pool as ref_fp
        from . import fixed_pool
        ref_fp.FixedPoolChunker = fixed_pool.FixedPoolChunker
        ref_models.FixedPoolChunker = fixed_pool.FixedPoolChunker
    except Exception:
        pass
    import dcasr.models.mamba_block as ref_mb
    for name in ("reverse_sequences", "MambaBlock", "MambaStack"):
        setattr(ref_mb, name, getattr(mamba_block, name))
    import dcasr.models.encoder as ref_enc
    for name in ("DCASREncoder", "EncoderOutput", "ConvSubsampling4", "build_chunker"):
        setattr(ref_enc, name, getattr(encoder, name))
    try:                                                        # the step after the path: CTC head (projection + fused loss)
        import dcasr.decoders.ctc as ref_ctc
        from . import ctc
        ref_ctc.CTCHead = ctc.CTCHead
    except Exception:
        ref_ctc = None
    if "dcasr.tasks.asr_task" in sys.modules:
        sys.modules["dcasr.tasks.asr_task"].DCASREncoder = encoder.DCASREncoder
        if ref_ctc is not None:
            sys.modules["dcasr.tasks.asr_task"].CTCHead = ref_ctc.CTCHead
    if patch_trainer:
        try:                                                    # needs the trainer's own third-party imports (editdistance)
            import dcasr.training.trainer as ref_trainer
            from . import trainer_sync
            trainer_sync.patch_trainer(ref_trainer.Trainer)     # per-micro-batch OOM flag: host-side collective, no .item()
        except Exception:
            pass
