"""Host-side arithmetic of bench.py that must hold on a box this suite never sees (8 ranks sharing one host)."""
import bench

GIB = 1 << 30
SAMPLE = 2 * 3 * 1080 * 1920 * 4  # source + destination bytes of one sample in pinned memory


def test_whole_batch_is_pinned_when_the_host_can_spare_it():
    assert bench.host_ring_samples(256, 16, available_bytes=2000 * GIB, local_ranks=8) == 256
    assert bench.host_ring_samples(256, 16, available_bytes=3 * 256 * SAMPLE, local_ranks=1) == 256


def test_ring_is_whole_chunks_within_a_third_of_the_ranks_share():
    for avail, ranks in ((200 * GIB, 8), (64 * GIB, 8), (20 * GIB, 4), (1 * GIB, 8)):
        n = bench.host_ring_samples(256, 16, available_bytes=avail, local_ranks=ranks)
        assert n % 16 == 0 and 16 <= n < 256
        assert n == 16 or n * SAMPLE * 3 * ranks <= avail


def test_small_batches_are_never_cut():
    assert bench.host_ring_samples(8, 16, available_bytes=GIB, local_ranks=8) == 8
