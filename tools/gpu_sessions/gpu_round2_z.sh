#!/bin/bash
# composed patch linears + three-CTA scale attention: parity suite, A/B of the fused patch stage, bench
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests/test_parity_gpu.py -q > gpurun_out/z2_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/z2_parity.log
timeout -s KILL 600 python - > gpurun_out/z2_ab_patch.json 2> gpurun_out/z2_ab_patch.err <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import duoformer_tcga_b200 as duo
torch.manual_seed(0)
m = duo.build_model_no_extra_params(pretrained=False, depth=12, embed_dim=768, num_heads=12, num_classes=10, num_layers=4, proj_dim=768).cuda().eval()
x = torch.randn(256, 3, 224, 224, device='cuda')
vt = m.vision_transformer
res = {True: [], False: []}
with torch.no_grad():
    for f in res:
        vt.fuse_patch_linears = f
        for _ in range(3): m(x)
    torch.cuda.synchronize()
    for rnd in range(4):
        for f in res:
            vt.fuse_patch_linears = f
            m(x)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): m(x)
            e1.record(); torch.cuda.synchronize()
            res[f].append(round(e0.elapsed_time(e1) / 4, 2))
print(json.dumps({"batch": 256, "ms_per_forward": {("fused" if k else "sequential"): v for k, v in res.items()},
                  "median": {("fused" if k else "sequential"): sorted(v)[len(v) // 2] for k, v in res.items()}}))
PY
echo "ab rc=$?"; cat gpurun_out/z2_ab_patch.json
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/z2_bench.json 2> gpurun_out/z2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/z2_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['by_shape_NxK_epi'].items():
    if v['launches']>6: print(k.ljust(24), round(v['tflops'],1), round(v['ms_per_launch'],3), round(v['ms_per_step'],2))
for k,v in d['roofline_hbm']['kernels'].items(): print(k, round(v['frac'],3), round(v['ms_per_step'],2))
print(d.get('trunk_convs'))
PY
