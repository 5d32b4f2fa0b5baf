#!/usr/bin/env python
"""Secondary measurement at BASELINE.json configs[4] geometry: 64-rig-frame batches of 5 x 1280x1024, 2000 features / camera,
device-resident input, CUDA events on the launching stream (not the headline: bench.py measures configs[2])."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmavis_b200 as ob  # noqa: E402
from openmavis_b200.synth import synth_image  # noqa: E402

W, H, NCAM, NF, FRAMES = 1280, 1024, 5, 2000, 64
n_img = FRAMES * NCAM
base = np.stack([synth_image(W, H, 5000 + i) for i in range(10)])
imgs = torch.from_numpy(base[np.arange(n_img) % 10].copy()).cuda()
for i in range(n_img):                                       # decorrelate the copies
    imgs[i] = torch.roll(imgs[i], shifts=(i // 10) * 7, dims=1)
params = ob.make_params(n_img, NF, 20, 7, (0, 0))
rig = ob.RigExtractor(W, H, n_img, NF)
stream = torch.cuda.Stream()
rig.set_stream(stream.cuda_stream)
rig.set_profiling(True)
with torch.cuda.stream(stream):
    for _ in range(3):
        rig.extract_device(imgs.data_ptr(), n_img, params)
    stream.synchronize()
    stage = rig.stage_ms()
    rig.set_profiling(False)
    rig.extract_device(imgs.data_ptr(), n_img, params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        rig.extract_device(imgs.data_ptr(), n_img, params)
    e1.record(stream)
    stream.synchronize()
ms = e0.elapsed_time(e1) / 10
res = rig.fetch(n_img)
print(json.dumps({'workload': 'sweep_64x5x1280x1024_nf2000', 'ms_per_step': ms, 'rig_frames_per_s': FRAMES / (ms * 1e-3),
                  'megapixels_per_s': n_img * W * H / (ms * 1e-3) / 1e6, 'keypoints_per_image': float(np.mean([len(r[1]) for r in res])),
                  'stage_ms': stage}))
