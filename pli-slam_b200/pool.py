"""Host-side sharding of camera streams over GPUs (replicas only — every stereo pair is independent, so there
is no data-path collective).  One process per GPU (rank = GPU): `shard_streams` says which streams a rank owns, `DevicePool`
keeps the contexts of one rank (one per CUDA stream kept in flight) and hands the calls of a stream always to the same one.
bench.py uses both; `ORB_SLAM3::plf_stream_owner` in pli-slam_b200/host/plf_frontend.hpp is the C++ twin."""
from .binding import Frontend


def shard_streams(n_streams, world_size, rank):
    """Stream s -> rank s mod world_size.  Returns the stream ids this rank owns."""
    return list(range(rank, n_streams, world_size))


def stream_owner(stream, world_size):
    return stream % world_size


def stream_seed(stream, frame):
    """Seed convention of SURVEY §8d for the synthetic streams: stream s, frame f -> 10000 * (s + 1) + f."""
    return 10_000 * (stream + 1) + frame


class DevicePool:
    """The contexts one rank keeps in flight on its GPU: `contexts` Frontends (one CUDA stream each) sized for calls of
    `streams_per_context` x `frames` pairs.  context_of(stream) is stable, so the frames of a stream stay in order."""

    def __init__(self, lib, device, contexts, streams_per_context, frames, **params):
        self.streams_per_context, self.frames = streams_per_context, frames
        self.ctx = [Frontend(lib, device=device, max_batch=streams_per_context * frames, **params) for _ in range(contexts)]

    def context_of(self, local_stream):
        return self.ctx[(local_stream // self.streams_per_context) % len(self.ctx)]

    def __iter__(self):
        return iter(self.ctx)

    def __len__(self):
        return len(self.ctx)

    def close(self):
        for f in self.ctx:
            f.close()
        self.ctx = []
