"""Run under torchrun (one rank per GPU): every rank records the same string methods, each level's PBS jobs are
sharded over the ranks, results reach every arena either by P2P stores from the kernel epilogue + flag barrier
(default) or by NCCL all-gather; every rank decrypts and checks against the plaintext oracle."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fhestring_b200.fhestring import FheSplit, MyClientKey
    from oracle import fhestring_plain as P
    from strcases import encode_args

    ck = MyClientKey.from_params(seed=31)          # same seed on every rank: same keys, same ciphertexts
    for exchange in ("p2p", "nccl"):
        sk = ck.get_server_key(device=local, arena_blocks=1 << 16, rank=rank, world=world)
        if exchange == "p2p":
            sk.engine.peer_attach(rank, world)
        else:
            sk.engine.comm_init(rank, world)
        pp = ck.get_public_parameters()
        s, pat = "the quick brown fox jumps over the lazy dog", "lazy"
        S, Pt = ck.encrypt(s, 2, pp, sk.key), ck.encrypt_no_padding(pat)
        assert ck.decrypt_char(sk.contains(S, Pt, pp)) == 1
        assert ck.decrypt_char(sk.find(S, Pt, pp)) == s.find(pat)
        sk.reset()
        S = ck.encrypt(s, 2, pp, sk.key)
        got = ck.decrypt_padded(sk.replace(S, ck.encrypt_no_padding("quick"), ck.encrypt_no_padding("slow"), pp))
        enc = encode_args("replace", [s, "quick", "slow"], 2)
        assert got == P.replace(*enc), (rank, exchange)
        sk.reset()
        bufs, found = FheSplit.decrypt(sk.split(ck.encrypt(" Mary had a", 1, pp, sk.key), ck.encrypt_no_padding(" "), pp), ck)
        assert [b for b in bufs if b] == ["Mary", "had", "a"] and found == 1
        if exchange == "p2p":
            assert not sk.engine.peer_timed_out()
        torch.cuda.synchronize()
        dist.barrier()
        sk.engine.close()
        if rank == 0:
            print(f"multi-gpu ok: world {world}, exchange {exchange}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
