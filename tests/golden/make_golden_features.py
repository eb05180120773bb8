"""Golden vectors for the beatmap feature builder from the UNMODIFIED reference
``data_loading.py`` (calc_distances :146-151, split_and_process_sequence_no_augment :172-187,
window_and_relative_time :195-203).  ``data_loading`` imports the third-party ``slider`` package, which is
absent here and is only used for .osu parsing: a stub module is injected so that the arithmetic
functions can be imported unchanged.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_features.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OSU_DIFFUSION_REF", "/root/reference")

for name in ("slider", "slider.beatmap", "slider.position", "slider.curve", "slider.mod"):
    sys.modules.setdefault(name, types.ModuleType(name))
for mod, names in (("slider", ["Beatmap", "Position"]),
                   ("slider.beatmap", ["Beatmap", "HitObject", "Slider", "Spinner", "Circle", "HoldNote", "TimingPoint"]),
                   ("slider.position", ["Position"]),
                   ("slider.curve", ["Linear", "Catmull", "Perfect", "MultiBezier", "Curve"]),
                   ("slider.mod", ["circle_radius"])):
    for n in names:
        setattr(sys.modules[mod], n, type(n, (), {}))
sys.path.insert(0, REF)
import data_loading as ref  # noqa: E402

g = torch.Generator().manual_seed(17)
T = 300
seq = torch.zeros(19, T)
seq[0] = torch.rand(T, generator=g) * 512
seq[1] = torch.rand(T, generator=g) * 384
seq[2] = 12345.0 + torch.cumsum(torch.randint(20, 900, (T,), generator=g).float(), 0)
kind = torch.randint(0, 16, (T,), generator=g)
seq[3:] = torch.nn.functional.one_hot(kind, 16).float().t()
seq[0, 5], seq[1, 5] = seq[0, 4], seq[1, 4]  # a zero distance (stacked objects)

dist = ref.calc_distances(seq.clone())
(x, o, c), n = ref.split_and_process_sequence_no_augment(seq.clone())
assert n == T
s, e, shift = 40, 168, 54321.5
xw, ow, cw = x[:, s:e], o[s:e] - o[s] + shift, c[:, s:e]  # window_and_relative_time with random.random()*1e5 = shift
np.savez_compressed(os.path.join(HERE, "features.npz"), seq=seq.numpy(), dist=dist.numpy(), x=x.numpy(), o=o.numpy(),
                    c=c.numpy(), o_sampling=(o - o[0]).numpy(), win=np.array([s, e]), shift=np.float32(shift),
                    xw=xw.numpy(), ow=ow.numpy(), cw=cw.numpy())
print("features.npz", os.path.getsize(os.path.join(HERE, "features.npz")) // 1024, "KiB")
