"""dcol_b200 — B200-native batched drop-in for DifferentiableCollisions.jl's proximity path.

Import as `import dcol_b200 as dc` (repo-root shim) — the directory name carries a dot and is
not importable directly.
"""
from .primitives import (  # noqa: F401
    ATT_MRP, ATT_QUAT, AbstractPrimitive, AbstractPrimitiveMRP,
    Polytope, Capsule, Cylinder, Cone, Sphere, Polygon, Ellipsoid,
    PolytopeMRP, CapsuleMRP, CylinderMRP, ConeMRP, SphereMRP, PolygonMRP, EllipsoidMRP,
    create_rect_prism, create_n_sided, mass_properties, mrp_from_q, q_from_mrp, dcm_from_q, dcm_from_mrp,
)
from .shapes import ShapeTable  # noqa: F401
from .api import (  # noqa: F401
    Context, Group, default_context, proximity, proximity_gradient, proximity_jacobian, MEM_HOST, MEM_DEVICE,
)
from .problem_matrices import problem_matrices, combine_problem_matrices, pair_problem  # noqa: F401
from . import flops_model  # noqa: F401
from .sharding import shard, shard_bounds, gather_rows, scene_wrench_sharded, bind_to_gpu_numa_node, bind_disjoint_cpus  # noqa: F401
