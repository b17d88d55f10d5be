"""`src`-shaped surface: the module layout and entry points of the reference's `src/` package
(config, dataclass, fileds, camera, scene, sdf, pathtracer, renderer, postprocessor, ibl) over
the CUDA hot path.  Unlike the reference, importing these modules has no side effects: the GPU
context is created by the first kernel call (refresh / pathtrace / post_process)."""
