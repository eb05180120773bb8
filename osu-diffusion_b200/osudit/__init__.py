"""Host-side glue of the B200-native osu-diffusion hot path: ctypes binding (`_lib`), tensor-level
wrappers (`ops`), the per-forward launch schedule (`engine`) and synthetic inputs (`synth`)."""
