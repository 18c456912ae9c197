"""Regenerate docs/sass/: full SASS of the two headline kernels and a per-kernel table of the
Blackwell-specific mnemonics (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA, LDGMC =
multimem.ld_reduce, ...).  Runs on a GPU-less host: cuobjdump only reads the objects."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "baton_b200", "csrc", "build")
OUT = os.path.join(ROOT, "docs", "sass")
KEY = re.compile(r"\b(STTM(?:\.[A-Za-z0-9_]+)*|UTC[A-Z0-9]+(?:\.[A-Z0-9_]+)*|UTMA[A-Z]+(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Za-z0-9_]+)*|STTM|UBLKCP|"
                 r"SYNCS(?:\.[A-Z0-9_]+)*|LDGMC(?:\.[A-Z0-9_]+)*|UCGABAR[A-Z_.]*|MAPA(?:\.[A-Z0-9_]+)*|"
                 r"LDS(?:\.[A-Z0-9_]+)*|RED(?:\.[A-Z0-9_]+)+|[A-Z]+\.E\.[0-9A-Z.]*SYS|HMMA[A-Z0-9_.]*|ACQBULK|"
                 r"ATOMG?(?:\.[A-Z0-9_]+)*|ELECT|CCTL[A-Z.]*|MEMBAR[A-Z.]*)")
os.makedirs(OUT, exist_ok=True)
lines = ["# SASS evidence per kernel (`cuobjdump -sass`, sm_100a)", "",
         "`UTCHMMA` = tcgen05.mma kind::f16, `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tiled",
         "load, `SYNCS.*` = mbarrier, `LDGMC` = multimem.ld_reduce (NVLS in-switch reduction), `UCGABAR` = cluster",
         "barrier, `*.SYS` = system-scope (cross-GPU) loads/stores, `UTCHMMA.2CTA` / `UTCBAR.2CTA.MULTICAST` = cta_group::2",
         "MMA and its multicast commit, `UTCQMMA` = block-scaled (MXFP8) MMA, `UTCCP` = tcgen05.cp (scale factors",
         "into TMEM), `UTMALDG.4D.IM2COL` = TMA im2col-mode load (implicit-GEMM convolution operands).",
         "`multimem.st` has NO mnemonic of its own: it is emitted as `STG.E.128.STRONG.SYS` whose address register is the",
         "multicast mapping -- the same register the preceding `LDGMC.E.HPADD.BF16x8` (multimem.ld_reduce) reads through",
         "(fedavg.sass: `LDGMC ... [R60.64]` followed by `STG.E.128.STRONG.SYS desc[..][R60.64]`); the replication is a",
         "property of the address.  Full listings of the headline instantiations: `gemm_tcgen05.sass` (incl. the",
         "implicit-GEMM forward / wgrad / dgrad modes), `gemm_fp8.sass`, `fedavg.sass`, `attention.sass`, `norm.sass`", "(cluster BatchNorm backward).", ""]
# full listings only for the headline instantiations (the complete objects are > 10 MB of text)
FULL = {"gemm_tcgen05": ("gemm_bf16_2cta_kernelILi256", "gemm_bf16_persistent_kernelILi256", "gemm_bf16_fixed_kernelILi256ELi4ELi0",
                         "gemm_bf16_tcgen05_kernelILi256ELb1ELi0", "gemm_bf16_fixed_kernelILi64ELi8ELi1",
                         "gemm_bf16_fixed_kernelILi64ELi8ELi2", "gemm_bf16_fixed_kernelILi64ELi8ELi3"),
        "attention": ("attention_fwd", "attention_bwd"),
        "norm": ("bn_bwd_cluster_kernelILi1",),
        "fedavg": ("fedavg_allreduce_kernelILi1", "fedavg_allreduce_kernelILi2"),
        "gemm_fp8": ("gemm_fp8_kernelILi128",)}
for f in ["gemm_tcgen05", "gemm_fp8", "quant", "attention", "im2col_tma", "fedavg", "elementwise", "conv", "norm", "loss",
          "gemm_simt"]:
    obj = os.path.join(BUILD, f + ".o")
    if not os.path.exists(obj):
        continue
    txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    lines.append("## {}.cu".format(f))
    full = []
    for fn in re.split(r"\n\s*Function : ", txt)[1:]:
        name = fn.split("\n", 1)[0].strip()
        if any(k in name for k in FULL.get(f, ())):
            full.append("\tFunction : " + fn)
        n_instr = len(re.findall(r"/\*[0-9a-f]{4,6}\*/", fn))
        c = collections.Counter(m.group(1) for m in KEY.finditer(fn))
        demangled = subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()
        lines.append("- `{}` ({} instructions)".format(demangled[:120], n_instr))
        if c:
            lines.append("    " + ", ".join("{} x{}".format(k, v) for k, v in sorted(c.items())))
    if full:
        # keep the instruction text, drop the second (encoding-only) line of every instruction and the trailing encoding
        # column: halves the listing without losing a mnemonic
        body = "\n".join(full)
        body = re.sub(r"\n\s*/\* 0x[0-9a-f]{16} \*/\s*(?=\n)", "", body)
        body = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/", "", body)
        open(os.path.join(OUT, f + ".sass"), "w").write(body)
    lines.append("")
open(os.path.join(OUT, "MNEMONICS.md"), "w").write("\n".join(lines))
print("wrote", OUT)
