from ._core import tree_map as map, tree_leaves as leaves  # noqa: F401
