"""The YtY summation classes (csrc/als_kernels.cuh GramMap, pio_als.cu gram_layout) restated in Python: class g = the
degree-rank positions p with p mod 16 in {g, 15 - g}.  The device code relies on two facts checked here for every world
size that divides 8: (1) all rows of a class are owned by ONE rank under assign_internal_kernel's serpentine dealing, and
that rank is the one gram_layout names; (2) the slots of a rank's classes are contiguous and in rank order, so an
all-gather of the slot sums concatenates correctly.  No GPU needed."""
import numpy as np
import pytest

GROUPS = 8


def owner_of_position(p, w):
    blk, pos = divmod(p, w)
    return (w - 1 - pos) if (blk & 1) else pos


def gram_layout(w, me):
    my_groups = GROUPS // w
    seen = [0] * w
    slot_of, cls = [0] * GROUPS, []
    for g in range(GROUPS):
        owner = owner_of_position(g, w)
        slot_of[g] = owner * my_groups + seen[owner]
        seen[owner] += 1
        if owner == me:
            cls.append(g)
    return my_groups, me * my_groups, cls, slot_of


def class_positions(g, n_rows):
    rem = n_rows % 16
    t_count = 2 * (n_rows // 16) + (1 if g < rem else 0) + (1 if 15 - g < rem else 0)
    return [16 * (t >> 1) + ((15 - g) if (t & 1) else g) for t in range(t_count)]


@pytest.mark.parametrize("w", [1, 2, 4, 8])
@pytest.mark.parametrize("n_rows", [0, 1, 7, 15, 16, 17, 31, 100, 1000, 4097])
def test_classes_are_rank_local_and_cover_every_row_once(w, n_rows):
    seen = np.zeros(n_rows, np.int32)
    for g in range(GROUPS):
        pos = class_positions(g, n_rows)
        assert pos == sorted(pos) and all(0 <= p < n_rows for p in pos)
        for p in pos:
            seen[p] += 1
        owners = {owner_of_position(p, w) for p in pos}
        assert len(owners) <= 1
        if owners:
            assert owners.pop() == owner_of_position(g, w)        # gram_layout derives the owner from position g
    assert (seen == 1).all()


@pytest.mark.parametrize("w", [1, 2, 4, 8])
def test_slots_are_contiguous_per_rank(w):
    all_slots = []
    for me in range(w):
        my_groups, slot0, cls, slot_of = gram_layout(w, me)
        assert len(cls) == my_groups
        assert [slot_of[g] for g in cls] == list(range(slot0, slot0 + my_groups))     # launch-local group lg -> slot0 + lg
        all_slots += [slot_of[g] for g in cls]
    assert sorted(all_slots) == list(range(GROUPS))
