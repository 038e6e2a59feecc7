#!/usr/bin/env python
"""sass_excerpt.py: instruction-mix table and hot-loop excerpts of the shipped library's kernels (cuobjdump -sass), written to
profiles/r02_sass_<kernel>.txt — the evidence behind DESIGN.md's statements about packed FP32 (FFMA2 / FADD2 / FMUL2), tensor
memory (LDTM / STTM), the TMA store (UTMASTG), cp.async (LDGSTS) and REDUX in the fused kernels."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mlx_audio_plus_b200", "lib", "libb200audio.so")
OUT = os.path.join(ROOT, "profiles")
KERNELS = {
    "k1_tma_whisper128": r"fast_logmel_tma_kernel.*MelSpec_whisper128E",
    "k1_512_parakeet80": r"fast_logmel_kernel.*Li16ELi16ELi160.*Lb1ELb1E.*MelSpec_parakeet80E",
    "k1_1024_vocos100": r"fast_logmel_kernel.*Li32ELi16ELi256.*MelSpec_vocos100E",
    "k3_fast_istft": r"fast_istft_kernel.*Lb0EEE",
    "k4_istft_small_polar": r"istft_small_kernelILi20ELb1EEE",
    "k4_stft_small_20": r"stft_small_kernelILi20EEE",
    "k2_generic_forward": r"frontend_generic_kernel",
    "k3b_generic_inverse": r"istft_generic_kernel",
}
INTEREST = ("FFMA2", "FADD2", "FMUL2", "LDTM", "STTM", "UTMASTG", "UTMALDG", "LDGSTS", "REDUX", "MUFU", "UTCBAR", "SYNCS")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s+Function : ", sass)
    for tag, rx in KERNELS.items():
        body = next((f for f in funcs[1:] if re.match(rx if rx.startswith("^") else ".*" + rx, f.split("\n", 1)[0])), None)
        if body is None:
            print("not found:", tag)
            continue
        name = body.split("\n", 1)[0]
        lines = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        ops = []
        for l in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
            if m:
                ops.append(m.group(1))
        cnt = collections.Counter(ops)
        with open(os.path.join(OUT, f"r02_sass_{tag}.txt"), "w") as f:
            f.write(f"{name}\n{len(ops)} SASS instructions (static)\n\ninstruction mix (static, top 30):\n")
            for k, v in cnt.most_common(30):
                f.write(f"  {k:10s} {v}\n")
            f.write("\nBlackwell-specific / notable:\n")
            for k in INTEREST:
                n = sum(v for kk, v in cnt.items() if kk.startswith(k))
                f.write(f"  {k:10s} {n}\n")
            # excerpts: 12 lines around the first occurrence of each notable mnemonic
            for k in ("UTMASTG", "LDTM", "FFMA2", "REDUX", "LDGSTS"):
                idx = next((i for i, l in enumerate(lines) if re.search(r"\b" + k, l)), None)
                if idx is None:
                    continue
                f.write(f"\n--- around the first {k} ---\n")
                for l in lines[max(0, idx - 6): idx + 7]:
                    f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() + "\n")
        print(tag, len(ops), {k: sum(v for kk, v in cnt.items() if kk.startswith(k)) for k in INTEREST})


if __name__ == "__main__":
    main()
