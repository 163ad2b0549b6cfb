"""Host-side batch dispatch over several B200s (one process per GPU, no collective on the data path).

Transport blocks are independent and so are cells; the only state that ties work to a GPU is the HARQ soft buffer of a
(cell, rnti, harq_id), which lives in that GPU's HBM until the TB is acknowledged or the buffer expires
(rx_buffer_pool_impl.cpp:36-142). The dispatcher is therefore a pure function every rank evaluates identically - no
communication is needed to agree on the owner of a TB:

  * a NEW transmission is assigned by a stable hash of its HARQ key (or, with `balance=True`, to the least-loaded rank,
    load = code blocks assigned in the current slot, ties to the lowest rank);
  * a RETRANSMISSION goes where its HARQ process already lives (sticky), until `release()`.

Reference analogue: reserve() keyed by trx_buffer_identifier{rnti, harq_id} in rx_buffer_pool_impl.cpp:36-105.
"""
import hashlib
from dataclasses import dataclass


@dataclass(frozen=True)
class HarqKey:
    cell: int
    rnti: int
    harq_id: int

    def stable_hash(self) -> int:
        d = hashlib.blake2b(f"{self.cell}:{self.rnti}:{self.harq_id}".encode(), digest_size=8).digest()
        return int.from_bytes(d, "little")


class TbDispatcher:
    """Deterministic owner-of-TB map; every rank holds an identical copy and feeds it the same slot descriptions."""

    def __init__(self, world_size: int, balance: bool = False):
        assert world_size >= 1
        self.world_size = world_size
        self.balance = balance
        self.owner = {}          # HarqKey -> rank, while the HARQ process is active
        self.slot_load = [0] * world_size

    def begin_slot(self):
        self.slot_load = [0] * self.world_size

    def assign(self, key: HarqKey, new_data: bool, nof_codeblocks: int = 1) -> int:
        if not new_data and key in self.owner:
            rank = self.owner[key]
        elif self.balance:
            rank = min(range(self.world_size), key=lambda r: (self.slot_load[r], r))
        else:
            rank = key.stable_hash() % self.world_size
        self.owner[key] = rank
        self.slot_load[rank] += nof_codeblocks
        return rank

    def release(self, key: HarqKey):
        """TB acknowledged (CRC ok) or its buffer expired: the next transmission on this key is free to move."""
        self.owner.pop(key, None)

    def shard(self, tbs, rank: int):
        """`tbs`: iterable of (HarqKey, new_data, nof_codeblocks, payload...). Returns the entries this rank owns."""
        mine = []
        for tb in tbs:
            if self.assign(tb[0], tb[1], tb[2]) == rank:
                mine.append(tb)
        return mine
