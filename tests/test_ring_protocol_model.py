"""Executable statement of the mbarrier-parity argument behind DESIGN.md section 9 item 1 (no GPU involved).

An mbarrier phase wait names its phase by PARITY only: ``try_wait.parity(P)`` succeeds when the barrier's completed-phase
count c satisfies (c & 1) != P.  That is unambiguous as long as a waiter is never a whole phase ahead of the barrier.

* In every ring of this library but one, ALL consumer warps wait on EVERY stage use in order (fused DMMA kernel, wide
  GEMMs, k_stream<1,T>).  A warp reaches use u+1 of a stage only after it passed use u of the same stage, so c >= u+1
  when it waits: the parity it tests is the right one.
* k_stream_dmma hands each stage use to ONE warp, a different one per reuse.  The warp that owns use u+1 of a stage never
  waited on use u, so it may arrive while c == u (use u's TMA data still in flight) -- and parity (u+1)&1 equals the parity
  of the long-completed use u-1: the wait falls through on stale data.

The model below is the protocol reduced to its counters; the two tests show the fall-through for the single-owner ring
and its impossibility for the all-consumers ring, for the stage / warp counts the kernels use.
"""
import itertools


def try_wait(completed_phases: int, parity: int) -> bool:
    """PTX mbarrier.try_wait.parity semantics on a barrier whose phases complete one by one."""
    return (completed_phases & 1) != parity


def test_parity_wait_is_sound_when_a_waiter_has_seen_the_previous_use():
    # all-consumers ring: a warp waiting for use u+1 has passed use u, i.e. c in {u+1, u+2} (u+2 only after it arrived
    # itself, which it has not) -> c == u+1 means "not yet", and the wait can only succeed once c == u+2
    for u in range(0, 50):
        parity = (u + 1) & 1
        assert not try_wait(u + 1, parity)      # data of use u+1 not landed: must block
        assert try_wait(u + 2, parity)          # landed: must pass


def test_single_owner_ring_can_fall_through_one_phase_early():
    stages, warps = 14, 8                       # k_stream_dmma at d = 100: 14 stages, 8 consumer warps
    found = []
    for i in range(0, 200):                     # item i -> stage i % stages, use u = i // stages, owner warp i % warps
        s, u = i % stages, i // stages
        nxt = i + stages                        # the next use of the same stage
        owner_next = nxt % warps
        prev_item_of_owner = nxt - warps        # what that warp did just before
        # the owner of the next use is a different warp and its previous item uses a different stage: nothing in the
        # protocol makes it wait for use u of stage s
        assert owner_next != i % warps and prev_item_of_owner % stages != s
        # if item i's data has not landed (c == u) when that warp arrives, its parity test already succeeds
        if u >= 1 and try_wait(u, (u + 1) & 1):
            found.append((i, prev_item_of_owner))
    assert found, "expected the fall-through to exist"
    # the inversion it needs: the later item (i + stages - warps) landed while item i had not -- TMA completions are unordered
    assert all(prev - i == stages - warps for i, prev in found)


def test_no_wrap_means_no_reuse():
    # the routing rule of kernels_stream.cu (stream_dmma_ring_wraps): with at most `stages` items per CTA every stage is
    # used once (u == 0 only), and use 0 waits on parity 0 against c in {0, 1}: exact
    stages = 14
    for items in range(1, stages + 1):
        uses = {i // stages for i in range(items)}
        assert uses == {0}
    assert not try_wait(0, 0) and try_wait(1, 0)
    # sanity: parity pairs that alias (the root of the problem) are exactly two phases apart
    assert all(try_wait(c, p) == try_wait(c + 2, p) for c, p in itertools.product(range(6), (0, 1)))
