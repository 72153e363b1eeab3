"""B200-native TCAR train / full-catalog-eval hot path (drop-in for the reference's model_combine.py path).

The directory name mirrors the reference repo; import it under the alias ``tcar_b200`` through
``__graft_entry__.load_package()`` (the hyphens make a plain ``import`` impossible).
"""
__all__ = ["_native", "build", "params", "model_combine", "catalog_parallel", "parallel", "sampler", "device_sampler",
           "util", "synth", "main"]
