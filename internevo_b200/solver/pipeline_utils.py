"""Layer → pipeline-stage assignment.

``partition_uniform(L, P, C)`` gives every stage ``C`` virtual chunks; chunk ``c`` of stage ``p`` is a half-open block
range ``(start, end)`` inside the ``c``-th slab of ``L / C`` consecutive blocks.  Inside a slab the first
``P - (L / C) % P`` stages take ``floor`` blocks and the remaining stages one more, so later stages (which also carry the
head) are never lighter than earlier ones — the same assignment the reference produces
(``internlm/solver/pipeline_utils.py:9-34``), computed here from prefix sums instead of a running cursor.
"""
from typing import List, Tuple


def _slab_bounds(slab_len: int, stages: int) -> List[int]:
    """Prefix boundaries (``stages + 1`` values) of one slab of ``slab_len`` blocks."""
    small, extra = divmod(slab_len, stages)
    if small == 0:
        raise ValueError(f"{stages} pipeline stages cannot share {slab_len} layers: a stage would be empty")
    first_big = stages - extra          # stages [first_big, stages) hold small + 1 blocks
    return [p * small + max(0, p - first_big) for p in range(stages + 1)]


def partition_uniform(num_items: int, pipeline_parallel_size: int, num_chunks: int) -> List[List[Tuple[int, int]]]:
    if num_items % num_chunks:
        raise AssertionError(f"{num_items} layers do not split into {num_chunks} equal virtual chunks")
    slab = num_items // num_chunks
    cuts = _slab_bounds(slab, pipeline_parallel_size)
    layout = [[(c * slab + cuts[p], c * slab + cuts[p + 1]) for c in range(num_chunks)]
              for p in range(pipeline_parallel_size)]
    covered = sorted(i for stage in layout for a, b in stage for i in range(a, b))
    assert covered == list(range(num_items)), (layout, num_items)
    return layout
