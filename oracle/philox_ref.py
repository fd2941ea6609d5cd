"""TEST INFRASTRUCTURE ONLY (tests/ and __graft_entry__.smoke() may import this; the product path never does).

Host restatement of the counter-based sampling noise the CUDA path draws inside the vocabulary GEMM's epilogue
(`acvae_b200/csrc/gemm.cuh`: philox4x32_10, philox_uniform4), so that a test can reproduce the device's draws bit for bit
and feed them to the injected-noise path (which IS pinned to the reference: word_model.py:173-207).

The generator is Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11).  It is not
part of the reference (which calls torch.rand_like per step); parity of THIS file is pinned on the three known-answer vectors
of the Random123 distribution (kat_vectors: philox4x32 10), checked in tests/test_host_logic.py.

Layout: key = the 64-bit seed (low word, high word); counter = (sequence row, word group, decode step, call number); word w of
the vocabulary belongs to group (w // 128) * 32 + ((w % 128) // 64) * 16 + w % 16 and takes output word (w % 64) // 16;
u = (x >> 8) * 2**-24.  The device turns u into a Gumbel variate with the hardware logarithm (gumbel_from_u_fast), the oracle with
an exact one: ids can differ where two keys are closer than ~1e-6, which the test allows for (a fraction of a percent of rows).
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(key, ctr):
    """key: (k0, k1) ints; ctr: four uint32-valued arrays (broadcastable).  Returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(*ctr)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & _MASK, p0 & _MASK, n0, n2
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def sampling_uniforms(seed: int, call: int, steps: int, rows: int, vocab: int) -> np.ndarray:
    """u[steps, rows, vocab] float32: exactly what acvae_decode_sample draws for rng_state = {seed, call}."""
    groups = (vocab + 127) // 128 * 32
    g = np.arange(groups, dtype=np.uint64)[None, None, :]
    r = np.arange(rows, dtype=np.uint64)[None, :, None]
    t = np.arange(steps, dtype=np.uint64)[:, None, None]
    x = philox4x32_10((seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF), (r, g, t, np.uint64(call & 0xFFFFFFFF)))
    x = np.stack(x, axis=-1)                                   # [steps, rows, groups, 4]
    w = np.arange(vocab)
    u = x[:, :, (w // 128) * 32 + ((w % 128) // 64) * 16 + w % 16, (w % 64) // 16]
    return ((u >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
