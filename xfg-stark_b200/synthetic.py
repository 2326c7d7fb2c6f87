"""Synthetic burn-mint inputs of SURVEY.md §8(d): SplitMix64(seed "XFGSTARK" + proof_index) -> tx_prefix_hash[32],
recipient[20], secret[32]; network_id 4, target_chain_id 42161, version 1 (the values of src/benchmarks/mod.rs:442-444)."""
SEED = 0x584647535441524B   # "XFGSTARK"
_MASK = (1 << 64) - 1


def _splitmix64(state):
    while True:
        state = (state + 0x9E3779B97F4A7C15) & _MASK
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        yield z ^ (z >> 31)


def synthetic_inputs(index=0):
    g = _splitmix64(SEED + index)
    raw = b"".join(next(g).to_bytes(8, "little") for _ in range(11))
    return dict(burn=8_000_000, mint=8_000_000, tx_prefix_hash=raw[:32], recipient=raw[32:52], secret=raw[56:88],
                network_id=4, target_chain_id=42161, version=1)
