"""Development aid: prints the actual bf16-path errors vs the fp32 oracle (the asserted bounds live in test_gpu_parity.py)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_gpu_parity as t
O, _, _ = t._oracle()
for base, kw in (('small6', {}), ('full', dict(num_layers=2))):
    cfg, params, feats, metas, _ = t.make_case(base, **kw)
    with torch.no_grad():
        want = O.head_forward(params, cfg, feats, metas)
    for tc in (False, True):
        eng = t.engine_for(cfg, params, metas, 'bf16', tc=tc)
        out = eng.forward([f[0].to(t.DEV) for f in feats])
        bev, occ, flow = t.to_ref_layout({k: v.cpu() for k, v in out.items()}, cfg)
        agree = (out['occ_cls'].cpu().long() == want['occ'].softmax(-1).argmax(-1)[0]).float().mean().item()
        print(f'{base} tc={tc}: bev max {(bev - want["bev_embed"]).abs().max():.4f} mean {(bev - want["bev_embed"]).abs().mean():.5f} '
              f'occ max {(occ - want["occ"]).abs().max():.4f} flow max {(flow - want["flow"]).abs().max():.4f} argmax agree {agree:.4f}')
