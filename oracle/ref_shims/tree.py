"""Minimal stand-in for dm-tree (requirements.txt:14) used by /root/reference/main.py:607,634 on flat dicts."""


def map_structure(fn, *structs):
    first = structs[0]
    if isinstance(first, dict):
        return {k: map_structure(fn, *[s[k] for s in structs]) for k in first}
    if isinstance(first, (list, tuple)):
        return type(first)(map_structure(fn, *xs) for xs in zip(*structs))
    return fn(*structs)
