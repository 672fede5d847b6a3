"""Multi-GPU sharding of one FASTQ stream (placeholder, filled in below)."""


class ShardedJob:
    @classmethod
    def synthetic(cls, *a, **k):
        raise NotImplementedError
