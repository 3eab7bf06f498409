import os, sys, torch
sys.path.insert(0, "/root/repo")
from lvt_b200.config.presets import preset
from lvt_b200.modeling import build_model
SMALL = ["MODEL.AUTOREGRESSIVE.VT.BLOCKS_E", ((1, 16, 16),) * 2, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_E", (8, 8),
         "MODEL.AUTOREGRESSIVE.VT.BLOCKS_D", ((1, 16, 16),) * 2, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_D", (8, 8)]
cfgv = preset("DSFVT", SMALL + ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", "/tmp/x"])
cfgv.freeze()
vt = build_model(cfgv); vt.train(False)
video = torch.randint(0, 512, (1, 4, 16, 16, 16)).cuda(); video[:, :, 15:] = 0
vt.model.sample_incremental = False
outs = {}
for graph in (True, "eager", False, "eager"):
    vt.sampler_graph = graph
    torch.manual_seed(0)
    o = vt.sample_video(video.clone(), temp=1e-10, n_prime=15).cpu()
    for k, v in outs.items():
        d = (o[0, :, 15] != v[0, :, 15])
        first = d.reshape(4, -1).any(0).nonzero()
        print(graph, "vs", k, "mismatch", d.float().mean().item(), "first pos", first[0].item() if len(first) else None)
    outs[str(graph) + str(len(outs))] = o
