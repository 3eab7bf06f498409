"""CPU, authoring container only: re-checks the oracle against the LIVE unmodified reference (imported from
/root/reference through oracle/ref_shim.py) instead of the committed fixtures.  Skipped wherever the reference
is absent (the GPU box, CI): there tests/test_oracle_golden.py pins the same functions to the golden vectors
this comparison produced."""
import numpy as np
import pytest
import torch

from oracle import lvt_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="needs /root/reference (authoring container)")
torch.set_num_threads(8)


def _load_into(module, weights):
    sd = module.state_dict()
    module.load_state_dict({**sd, **weights}, strict=True)


def test_vq_indices_and_distances_match_live_reference():
    ref_shim.install()
    from vidgen.modeling.vq.vq_utils import vq
    from oracle import vq as ovq
    g = torch.Generator().manual_seed(99)
    x = torch.randn((777, 64), generator=g) * 0.3
    cb = torch.randn((512, 64), generator=g) * 0.3
    want = vq(x, cb)
    assert np.array_equal(O.vq_indices(x, cb).numpy(), want.numpy())
    z = x.t().reshape(1, 64, 777, 1).contiguous()
    assert np.array_equal(ovq.vq_argmin_c(z, cb[None]).view(-1).numpy(), want.numpy())


def test_dsfvt_two_layer_loss_and_logits_match_live_reference():
    ref_shim.install()
    from vidgen.modeling.meta_arch import build_model
    from vidgen.utils.events import EventStorage
    layers, batch = 1, 2
    blocks, heads = str(tuple([(1, 16, 16)] * layers)), str(tuple([8] * layers))
    cfg = ref_shim.reference_cfg("configs/vt/DSFVT.yaml", [
        "MODEL.AUTOREGRESSIVE.VT.BLOCKS_E", blocks, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_E", heads,
        "MODEL.AUTOREGRESSIVE.VT.BLOCKS_D", blocks, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_D", heads])
    torch.manual_seed(0)
    model = build_model(cfg)
    ocfg = O.VTConfig(blocks_e=((1, 16, 16),) * layers, heads_e=(8,) * layers, blocks_d=((1, 16, 16),) * layers,
                      heads_d=(8,) * layers)
    weights = O.synth_weights(O.dsfvt_param_shapes(ocfg), seed=31)
    _load_into(model.model, weights)
    model.train()
    context, slc, slice_idx, ignore = O.synth_vt_batch(batch, seed=13, cfg=ocfg)
    data = [{"context": context[i], "slice": slc[i], "slice_idx": slice_idx[i], "ignore_mask": ignore[i]}
            for i in range(batch)]
    with EventStorage(0):
        want = model(data, mode="supervised")["loss_cross_entropy"]
    sd = {k: v.clone() for k, v in weights.items()}
    got = O.vt_supervised_loss(context, slc, slice_idx, ignore, sd, ocfg)
    assert np.allclose(got.item(), want.item(), rtol=1e-6)
    with torch.no_grad():
        lw = torch.stack(model.model(context, slc, slice_idx))
        lg = torch.stack(O.vt_logits(context, slc, slice_idx, sd, ocfg))
    assert torch.allclose(lg, lw, rtol=1e-4, atol=1e-5)
