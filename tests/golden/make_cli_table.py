"""Generates cli_table.json from the reference's ArgParse table (mcmc_eap_chain.jl:19-153).
Run in the build container only (/root/reference does not exist on the GPU box):
    python tests/golden/make_cli_table.py
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
JOBS = [("/root/reference/mcmc_eap_chain.jl", os.path.join(HERE, "cli_table.json")),
        # the clustering driver's table, mcmc_clustering_eap_chain.jl:19-153
        ("/root/reference/mcmc_clustering_eap_chain.jl", os.path.join(HERE, "cli_table_clustering.json")),
        # the 2-D tree's only driver, 2D/mcmc_clustering_eap_chain.jl:19-133
        ("/root/reference/2D/mcmc_clustering_eap_chain.jl", os.path.join(HERE, "cli_table_clustering_2d.json"))]

# each entry: one or two quoted option strings followed by indented key = value lines
pat = re.compile(r'^\s*((?:"-[^"]+"\s*,?\s*)+)\n((?:\s+\w+\s*=.*\n)+)', re.M)
for SRC, OUT in JOBS:
    text = open(SRC, encoding="utf-8").read()
    table = text[text.index("@add_arg_table"):text.index("pargs = parse_args(s)")]
    entries = []
    for m in pat.finditer(table):
        names = re.findall(r'"(-[^"]+)"', m.group(1))
        body = dict(re.findall(r'^\s+(\w+)\s*=\s*(.*?)\s*;?\s*$', m.group(2), re.M))
        long = [x for x in names if x.startswith("--")][0]
        short = [x for x in names if not x.startswith("--")]
        e = {"long": long, "short": short[0] if short else None,
             "arg_type": body.get("arg_type"), "default": body.get("default"),
             "action": body.get("action")}
        entries.append(e)
    json.dump(entries, open(OUT, "w"), indent=1)
    print(len(entries), "options ->", OUT)
