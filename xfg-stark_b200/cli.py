"""`generate` / `validate` of the reference CLI over the CUDA backend.

Mirrors `xfg-stark-cli generate -i <data package> -o <proof>` (src/bin/xfg-stark-cli.rs:438-564) and the data-package schema
and validation of src/proof_data_schema.rs:10-24, 154-218, 275-320: same JSON field names in and out (`StarkProof{proof_data:
Vec<u8>, public_inputs, metadata}` :43-68), same hex handling (:715-736), same argument packing for
`XfgBurnMintProver::prove_burn_mint` (:476-528: transaction hash -> 32-byte tx_prefix_hash, address -> 20 bytes, the secret
key's UTF-8 bytes zero-padded to 32, network_id parsed as u32 with default 1, target chain 42161, commitment version 1).

    python -m xfg_stark_b200.cli generate -i package.json -o proof.json [--trace-log2 6] [--extension quadratic]
    python -m xfg_stark_b200.cli validate -i package.json
"""
import argparse
import datetime
import json
import sys

VALID_AMOUNTS_XFG = (0.8, 800.0)


def hex_to_bytes(h):
    """src/bin/xfg-stark-cli.rs:715-723"""
    return bytes.fromhex(h[2:] if h.startswith("0x") else h)


def hex_to_u64(h):
    """src/bin/xfg-stark-cli.rs:725-736"""
    b = hex_to_bytes(h)
    if len(b) < 8:
        raise ValueError("Hex string too short for u64")
    return int.from_bytes(b[:8], "little")


def validate_package(pkg):
    """StarkProofDataPackage::validate (src/proof_data_schema.rs:275-320) -> (is_valid, errors, warnings)"""
    errors, warnings = [], []
    bt, rc, sc = pkg["burn_transaction"], pkg["recipient"], pkg["secret"]
    try:
        amount = float(bt["burn_amount_xfg"])
    except ValueError:
        amount = 0.0
    if amount not in VALID_AMOUNTS_XFG:
        errors.append(f"Burn amount must be exactly 0.8 XFG or 800.0 XFG, got {amount}")
    if bt["transaction_hash"].startswith("0x"):
        errors.append("Fuego transaction hash should not start with 0x")
    if not rc["ethereum_address"].startswith("0x") or len(rc["ethereum_address"]) != 42:
        errors.append("Ethereum address must be 0x-prefixed 40-character hex")
    if len(sc["secret_key"]) < 8:
        errors.append("Secret key must be at least 8 characters")
    if bt.get("block_height", 0) == 0:
        warnings.append("Block height is 0 - please verify this is correct")
    if bt.get("timestamp", 0) == 0:
        warnings.append("Timestamp is 0 - please verify this is correct")
    return not errors, errors, warnings


def prover_arguments(pkg):
    """The 8 arguments of prove_burn_mint exactly as generate_proof builds them (src/bin/xfg-stark-cli.rs:476-528)."""
    bt = pkg["burn_transaction"]
    hex_to_u64(bt["transaction_hash"])                                   # :477-478, only a format check there
    txh = hex_to_bytes(bt["transaction_hash"])[:32].ljust(32, b"\0")     # :491-499
    rcpt = hex_to_bytes(pkg["recipient"]["ethereum_address"])[:20].ljust(20, b"\0")   # :501-507
    secret = pkg["secret"]["secret_key"].encode()[:32].ljust(32, b"\0")  # :509-515
    try:
        network_id = int(bt["network_id"])                               # :518 parse::<u32>().unwrap_or(1)
        if not 0 <= network_id < 1 << 32:
            network_id = 1
    except ValueError:
        network_id = 1
    return dict(burn_amount=bt["burn_amount_atomic"], mint_amount=bt["burn_amount_atomic"], tx_prefix_hash=txh, recipient_address=rcpt, secret=secret,
                network_id=network_id, target_chain_id=42161, commitment_version=1)


def generate_proof(pkg, prover):
    """-> the reference's `StarkProof` JSON object (src/proof_data_schema.rs:43-68, built at src/bin/xfg-stark-cli.rs:536-551)."""
    ok, errors, _ = validate_package(pkg)
    if not ok:
        raise ValueError("Data package validation failed: " + "; ".join(errors))
    a = prover_arguments(pkg)
    proof = prover.prove_burn_mint(a["burn_amount"], a["mint_amount"], a["tx_prefix_hash"], a["recipient_address"], a["secret"],
                                   a["network_id"], a["target_chain_id"], a["commitment_version"])
    bt = pkg["burn_transaction"]
    return {
        "proof_data": list(proof),                                       # serde's Vec<u8> = JSON array of numbers
        "public_inputs": {"burn_amount": bt["burn_amount_atomic"], "mint_amount": bt["burn_amount_atomic"], "txn_hash": bt["transaction_hash"],
                          "recipient_hash": pkg["recipient"]["ethereum_address"], "state": 0},
        "metadata": {"version": "1.0.0", "created_at": datetime.datetime.now(datetime.timezone.utc).isoformat(),
                     "description": f"STARK proof for {bt['burn_amount_xfg']} XFG burn", "network": pkg["metadata"]["network"]},
    }


def main(argv=None):
    ap = argparse.ArgumentParser(prog="xfg-stark-cli (B200 backend)")
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("generate"); g.add_argument("-i", "--input", required=True); g.add_argument("-o", "--output", required=True)
    g.add_argument("--trace-log2", type=int, default=6, help="trace length 2^k (6 = the reference's 64 rows)")
    g.add_argument("--extension", choices=["none", "quadratic", "cubic"], default="none")
    # ProofOptions::new arguments (XfgBurnMintProver::with_options, src/burn_mint_prover.rs:44-49); defaults = the reference's (src/burn_mint_prover.rs:28-35)
    g.add_argument("--queries", type=int, default=42); g.add_argument("--blowup", type=int, default=8); g.add_argument("--grinding", type=int, default=4)
    g.add_argument("--folding", type=int, default=8); g.add_argument("--remainder-degree", type=int, default=31)
    v = sub.add_parser("validate"); v.add_argument("-i", "--input", required=True)
    args = ap.parse_args(argv)
    pkg = json.load(open(args.input))
    ok, errors, warnings = validate_package(pkg)
    for w in warnings:
        print("warning:", w)
    if not ok:
        for e in errors:
            print("error:", e, file=sys.stderr)
        return 1
    if args.cmd == "validate":
        print("Data package validated successfully")
        return 0
    from ._binding import FieldExtension, ProofOptions, XfgBurnMintProver
    ext = {"none": FieldExtension.NONE, "quadratic": FieldExtension.QUADRATIC, "cubic": FieldExtension.CUBIC}[args.extension]
    opts = ProofOptions(num_queries=args.queries, blowup_factor=args.blowup, grinding_factor=args.grinding, field_extension=ext,
                        fri_folding_factor=args.folding, fri_remainder_max_degree=args.remainder_degree)
    prover = XfgBurnMintProver.with_options(128, opts, trace_log2=args.trace_log2)
    out = generate_proof(pkg, prover)
    json.dump(out, open(args.output, "w"), indent=2)
    print(f"Proof size: {len(out['proof_data'])} bytes; saved to {args.output}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
