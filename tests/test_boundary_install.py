"""The drop-in boundary, executed: `dcasr_b200.install()` followed by the REFERENCE's own builders
(`dcasr.tasks.asr_task.build_encoder` / `build_model`, src/dcasr/tasks/asr_task.py:27-58,129-146) must yield the B200
classes with the reference's module tree, state_dict keys/shapes and optimiser hooks.  Needs /root/reference (this
container only: the GPU box does not have it) -- construction is host-side and needs no GPU.  `editdistance` and
`omegaconf` are not installed here: the subprocess stubs them, as a launcher on such a box would."""
import json
import os
import subprocess
import sys

import pytest

from _util import PKG_DIR, REPO

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_SRC, "dcasr")), reason="reference checkout not present")

CODE = r'''
import json, sys, types
sys.path[:0] = [%(ref)r, %(pkg)r, %(repo)r]
for name in ("editdistance", "omegaconf"):              # absent third-party imports of the reference's trainer / scripts
    m = types.ModuleType(name)
    m.eval = lambda a, b: 0
    m.OmegaConf = type("OmegaConf", (), {})
    sys.modules[name] = m
import dcasr_b200
dcasr_b200.install()
import torch
from dcasr.tasks.asr_task import build_encoder, build_model
import dcasr.models.encoder as ref_enc, dcasr.models.mamba_block as ref_mb, dcasr.models.hnet_chunk as ref_ch
out = {}
for arch, N, chunker in (("A", 2, "dynamic"), ("B", 4, "dynamic"), ("A", 2, "fixed"), ("A", 1, "dynamic")):
    cfg = {"encoder": "dcasr", "head": "ctc", "frontend_conf": {"n_mels": 80},
           "encoder_conf": {"d_outer": 64, "d_main": 128, "n_enc": 1, "n_main": 2, "n_dec": 1, "n_mid": 1, "arch_type": arch,
                            "hnet": {"compression_N": N, "chunker": chunker}},
           "model_conf": {"ctc_weight": 1.0, "aed_weight": 0.0}}
    enc = build_encoder(cfg)
    model = build_model(cfg, 501)
    sd = enc.state_dict()
    out[f"{arch}{N}{chunker}"] = {
        "encoder_class": type(enc).__module__ + "." + type(enc).__name__,
        "model_encoder_class": type(model.encoder).__module__,
        "block_class": type(enc.enc.layers[0]).__module__, "mixer_class": type(enc.enc.layers[0].fwd).__module__,
        "chunker_class": type(enc.chunk if arch == "A" else enc.chunk1).__module__,
        "keys": {k: list(v.shape) for k, v in sd.items()},
        "no_decay": sorted(n for n, p in enc.named_parameters() if getattr(p, "_no_weight_decay", False)),
        "router": sorted(n for n, p in enc.named_parameters() if "router" in n.split(".") and n.split(".")[-2] in ("W_q", "W_k")),
    }
from dcasr.training.trainer import Trainer
out["trainer_sync"] = Trainer._any_rank_oom.__doc__ or ""
out["rebinds"] = [ref_enc.DCASREncoder.__module__, ref_mb.MambaStack.__module__, ref_ch.DynamicChunker.__module__,
                  sys.modules["mamba_ssm"].Mamba2.__module__]
try:
    build_encoder({"encoder": "dcasr", "frontend_conf": {"n_mels": 80}, "encoder_conf": {"d_outer": 64, "d_main": 128, "n_enc": 1,
                   "n_main": 1, "n_dec": 1, "arch_type": "C"}})
    out["bad_arch"] = "no error"
except ValueError as e:
    out["bad_arch"] = "ValueError"
print("RESULT" + json.dumps(out))
'''


def test_reference_builders_construct_the_b200_classes():
    r = subprocess.run([sys.executable, "-c", CODE % {"ref": REF_SRC, "pkg": PKG_DIR, "repo": REPO}], capture_output=True,
                       text=True, timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(line[0][6:])
    assert all(m.startswith("dcasr_b200") for m in out["rebinds"]), out["rebinds"]
    assert out["bad_arch"] == "ValueError"
    assert "dcasr_b200.trainer_sync" in out["trainer_sync"]            # per-micro-batch OOM flag rebound to the host-side collective
    # the oracle mirrors the reference's module tree (tests/test_oracle_encoder.py pins it against the reference's own
    # encoder.py): identical key sets and shapes mean reference checkpoints load with strict=True
    from oracle.encoder_ref import EncoderRef
    for tag, (arch, N, chunker) in {"A2dynamic": ("A", 2, "dynamic"), "B4dynamic": ("B", 4, "dynamic"),
                                    "A2fixed": ("A", 2, "fixed"), "A1dynamic": ("A", 1, "dynamic")}.items():
        o = out[tag]
        assert o["encoder_class"] == "dcasr_b200.encoder.DCASREncoder" and o["model_encoder_class"].startswith("dcasr_b200")
        assert o["block_class"].startswith("dcasr_b200") and o["mixer_class"].startswith("dcasr_b200")
        assert o["chunker_class"].startswith("dcasr_b200")
        ref = EncoderRef(n_mels=80, d_outer=64, d_main=128, n_enc=1, n_main=2, n_dec=1, n_mid=1, arch_type=arch, N=N, chunker=chunker)
        want = {k: list(v.shape) for k, v in ref.state_dict().items()}
        assert o["keys"] == want, (tag, set(o["keys"]) ^ set(want))
        assert o["no_decay"] and all(n.split(".")[-1] in ("A_log", "D", "dt_bias") for n in o["no_decay"])
        n_mixers = sum(1 for k in want if k.endswith("A_log"))
        assert len(o["no_decay"]) == 3 * n_mixers                      # optimiser hook of training/trainer.py:153-154
        expect_routers = 0 if (N == 1 or chunker == "fixed") else (2 if arch == "A" else 4)
        assert len(o["router"]) == expect_routers, (tag, o["router"])    # found by name, training/trainer.py:139-141
