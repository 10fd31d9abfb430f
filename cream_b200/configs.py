"""Supernet geometries and search spaces shipped by the reference
(AutoFormer/experiments/supernet/supernet-{T,S,B}.yaml)."""

SUPERNETS = {
    "T": dict(embed_dim=256, depth=14, num_heads=4, mlp_ratio=4.0),
    "S": dict(embed_dim=448, depth=14, num_heads=7, mlp_ratio=4.0),
    "B": dict(embed_dim=640, depth=16, num_heads=10, mlp_ratio=4.0),
}

SEARCH_SPACE = {
    "T": dict(mlp_ratio=[3.5, 4.0], num_heads=[3, 4], depth=[12, 13, 14], embed_dim=[192, 216, 240]),
    "S": dict(mlp_ratio=[3.0, 3.5, 4.0], num_heads=[5, 6, 7], depth=[12, 13, 14], embed_dim=[320, 384, 448]),
    "B": dict(mlp_ratio=[3.0, 3.5, 4.0], num_heads=[9, 10], depth=[14, 15, 16], embed_dim=[528, 576, 624]),
}
