"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py` (frequency-domain engine) into the
per-iteration table under profiles/.   python scripts/launch_summary.py LAUNCHES.csv PLAIN.json OUT.md"""
import csv
import json
import sys

src, plain_json, out_md = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hdr, data = None, []
for r in rows:
    if hdr is None:
        if r and r[0] == "ID":
            hdr = r
        continue
    data.append(dict(zip(hdr, r)))
ks = [(d["Kernel Name"], float(d["Metric Value"].replace(",", "")), d["Grid Size"], d["Block Size"])
      for d in data if d.get("Metric Name") == "gpu__time_duration.sum"]
idx = [i for i, k in enumerate(ks) if "ifft_numW_kernel<float>" in k[0]]
a, b = idx[3] - 1, idx[4] - 1            # one iteration of the timed region: from its numW product to the next one
it = ks[a:b]
tot = sum(k[1] for k in it)
plain = json.load(open(plain_json))
km = plain["roofline"]["kernel_ms"]
seen = {}


def nth(key):
    seen[key] = seen.get(key, 0) + 1
    return seen[key] - 1


def role(n):
    if "tc_kernel<5>" in n:
        return ["numW: per-frequency conj(H^) X^ over the spectrum of X (streams the spectrum planes)", "Gram partial of the new H: conj(H^) H^"][min(nth("5"), 1)]
    if "tc_kernel<4>" in n:
        return ["numH: per-frequency conj(W^) X^ (streams the spectrum planes)", "denomH: per-frequency conj(C^) H^"][min(nth("4"), 1)]
    if "tc_kernel<3>" in n:
        return ["denomW = G W (plain tcgen05 GEMM)", "W W^T (plain tcgen05 GEMM)"][min(nth("3"), 1)]
    if "ifft_numW_kernel<float>" in n:
        return "inverse FFT -> numW"
    if "ifft_numW_kernel<double>" in n:
        return "inverse FFT -> Gram partial Rg (fp64)"
    if "ifft_numH" in n:
        return ["inverse FFT -> numH", "inverse FFT -> denomH"][min(nth("i"), 1)]
    if "fft_h_kernel" in n:
        return ["FFT of H blocks (hop B-2L+2, both halos) for denomH", "FFT of the owned H segments -> Ah (numW of the next iteration + Gram)",
                "FFT of full H blocks -> Hf (Gram)"][min(nth("h"), 2)]
    if "fft_w_kernel" in n:
        return ["FFT of W -> Aw", "FFT of the lag table C -> Ac"][min(nth("w"), 1)]
    for key, txt in (("build_G_split", "G = Htilde Htilde^T from Rg + tail, as bf16 planes (diagonal walk)"),
                     ("mu_update", "multiplicative update (mult.jl:37-38 / 51-52)"), ("split_W", "bf16 planes of W for the plain GEMMs"),
                     ("lag_table", "lag table C from W W^T"), ("denomH_tail", "truncated tail of denomH (prefix form)"),
                     ("s2_dot_G", "expansion loss: <W W^T, Htilde Htilde^T> (diagonal walk)"), ("dot_partial", "expansion loss: <numH, H>"),
                     ("h_tail", "last L-1 columns of H for the Gram tail"), ("reduce", "deterministic reduction")):
        if key in n:
            return txt
    return ""


out = ["# ncu launch list, frequency-domain engine (default), FULL c4 size (N=4096, T=4194304, K=64, L=100, fp32 MU, 1 GPU)\n",
       '    CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-calibrated"',
       f"    $CMD && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/{__import__('os').path.basename(src)} $CMD\n",
       f"Raw list: `profiles/{__import__('os').path.basename(src)}` ({len(ks)} launches).  One MU iteration with the expansion loss (the default of",
       "`bench.py`), taken from the timed region; per-launch times under ncu are cold-cache and serialised: compare shares.",
       f"The plain (un-profiled) run of the same command measured **{plain['ms_per_step']:.2f} ms/iteration** with CUDA events; its live",
       f"per-launch times of the two products: numW (`TC_FQC`) {km['corr']['total_ms'] / km['corr']['launches']:.2f} ms, numH (`TC_FQT`) "
       f"{km['transconv']['total_ms'] / km['transconv']['launches']:.2f} ms"]
prod = sum(k[1] for k in it if ("tc_kernel<4>" in k[0] or "tc_kernel<5>" in k[0]) and k[1] > 5e6) / 1e6
out.append(f"= {100 * (km['corr']['total_ms'] + km['transconv']['total_ms']) / (plain['ms_per_step'] * plain['steps']):.0f} % of the step live vs "
           f"{100 * prod / (tot / 1e6):.0f} % under ncu (the live step also carries launch gaps and the sustained power-capped clock).\n")
out += ["| ms | share | grid | block | kernel | role |", "|---:|---:|---|---|---|---|"]
for k in it:
    nm = k[0].split("(")[0].replace("void ", "")
    out.append(f"| {k[1] / 1e6:.3f} | {100 * k[1] / tot:.1f}% | {k[2]} | {k[3]} | `{nm}` | {role(k[0])} |")
out.append(f"| **{tot / 1e6:.3f}** | 100% | | | one iteration | |\n")
fft = sum(k[1] for k in it if "fft" in k[0]) / 1e6
out.append(f"Per-frequency products over the spectrum of X: {prod:.1f} ms ({100 * prod / (tot / 1e6):.0f} %); FFT kernels: {fft:.1f} ms "
           f"({100 * fft / (tot / 1e6):.0f} %); everything else (W-sized GEMMs, updates, tail, loss terms, the small products): "
           f"{tot / 1e6 - prod - fft:.1f} ms.  The pieces that do not depend on T (inverse FFT of numW, G build, G W, FFT of W, planes of W, "
           f"W W^T, denomH tail, loss terms) are the serial fraction of the T-sharded multi-GPU run.")
open(out_md, "w").write("\n".join(out) + "\n")
print("\n".join(out))
