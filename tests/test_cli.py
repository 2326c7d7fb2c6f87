"""The CLI `generate` path (src/bin/xfg-stark-cli.rs:438-564, src/proof_data_schema.rs): schema,
validation and argument packing on the CPU; the proof itself on the GPU, byte-equal to the oracle and accepted by its verifier."""
import json
import os

import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))


def package():
    return json.load(open(os.path.join(HERE, "golden", "data_package.json")))


def test_validation_and_argument_packing():
    from xfg_stark_b200 import cli
    pkg = package()
    ok, errors, warnings = cli.validate_package(pkg)
    assert ok and not errors and not warnings
    a = cli.prover_arguments(pkg)
    assert a["burn_amount"] == a["mint_amount"] == 8_000_000 and a["network_id"] == 4 and a["target_chain_id"] == 42161 and a["commitment_version"] == 1
    assert a["tx_prefix_hash"] == bytes.fromhex(pkg["burn_transaction"]["transaction_hash"]) and len(a["recipient_address"]) == 20
    assert a["secret"] == b"correct horse battery staple".ljust(32, b"\0")
    assert cli.hex_to_u64("0x0102030405060708ff") == int.from_bytes(bytes(range(1, 9)), "little")
    bad = package(); bad["burn_transaction"]["burn_amount_xfg"] = "1.5"; bad["burn_transaction"]["transaction_hash"] = "0xabc"
    bad["recipient"]["ethereum_address"] = "742d"; bad["secret"]["secret_key"] = "short"; bad["burn_transaction"]["block_height"] = 0
    ok, errors, warnings = cli.validate_package(bad)
    assert not ok and len(errors) == 4 and len(warnings) == 1
    pkg["burn_transaction"]["network_id"] = "fuego-mainnet"          # not a number: unwrap_or(1)
    assert cli.prover_arguments(pkg)["network_id"] == 1
    assert cli.main(["validate", "-i", os.path.join(HERE, "golden", "data_package.json")]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("ext", ["none", "quadratic"])
def test_generate_matches_oracle(tmp_path, ext):
    from xfg_stark_b200 import cli
    out = tmp_path / "proof.json"
    assert cli.main(["generate", "-i", os.path.join(HERE, "golden", "data_package.json"), "-o", str(out), "--trace-log2", "6", "--extension", ext]) == 0
    d = json.load(open(out))
    assert set(d) == {"proof_data", "public_inputs", "metadata"} and d["public_inputs"]["state"] == 0 and d["public_inputs"]["burn_amount"] == 8_000_000
    proof = bytes(d["proof_data"])
    a = cli.prover_arguments(package())
    pi, ac, _ = orc.pack_inputs(a["burn_amount"], a["mint_amount"], a["tx_prefix_hash"], a["recipient_address"], a["secret"], a["network_id"], a["target_chain_id"], a["commitment_version"])
    opts = (42, 8, 4, 2 if ext == "quadratic" else 1, 8, 31)
    assert proof == orc.prove(orc.build_trace(pi, ac, 64), pi, ac, opts)
    assert orc.verify(proof, pi, ac, opts) == ""


@pytest.mark.gpu
def test_generate_with_other_proof_options(tmp_path):
    """`with_options` through the CLI: cubic extension, blowup 16, folding 4"""
    from xfg_stark_b200 import cli
    out = tmp_path / "proof.json"
    assert cli.main(["generate", "-i", os.path.join(HERE, "golden", "data_package.json"), "-o", str(out), "--trace-log2", "7", "--extension", "cubic",
                     "--blowup", "16", "--folding", "4", "--queries", "30", "--grinding", "6", "--remainder-degree", "15"]) == 0
    proof = bytes(json.load(open(out))["proof_data"])
    a = cli.prover_arguments(package())
    pi, ac, _ = orc.pack_inputs(a["burn_amount"], a["mint_amount"], a["tx_prefix_hash"], a["recipient_address"], a["secret"], a["network_id"], a["target_chain_id"], a["commitment_version"])
    opts = (30, 16, 6, 3, 4, 15)
    assert proof == orc.prove(orc.build_trace(pi, ac, 128), pi, ac, opts)
    assert orc.verify(proof, pi, ac, opts) == ""
