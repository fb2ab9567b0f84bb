#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for D in 0 1 2 3; do
VOX_G5_DEBUG=$D timeout 300 python - <<PY
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
import voxtral_mini_realtime_rs_b200 as vx
from voxtral_mini_realtime_rs_b200 import synth
G="/dev/shm/voxtral_synth_s42.gguf"
if not os.path.exists(G): synth.write_synthetic_gguf(G, synth.VoxtralConfig(), seed=42)
m=vx.Q4ModelLoader.from_file(G).load(0,max_batch=8,max_mel_frames=2400)
mel=np.zeros((8,128,2344),np.float32)
import time
for _ in range(2): m.encode_audio(mel)
t=time.perf_counter()
for _ in range(5): m.encode_audio(mel)
print("VOX_G5_DEBUG=$D encode_audio B=8 (incl. H2D/D2H of mel/embeds): %.2f ms" % ((time.perf_counter()-t)/5*1e3))
PY
done
