"""`jax.tree_util` facade of jaxlite."""
from ._impl import (tree_flatten, tree_unflatten, tree_structure, tree_leaves, tree_map, tree_reduce,
                    register_pytree_node_class, PyTreeDef)
