"""``tmlauncher`` — command-line launcher (ref ``theanompi/bin/tmlauncher``, bash).

    tmlauncher -cfg=session.cfg
    tmlauncher -file=theanompi_b200.models.alex_net -class=AlexNet -r=BSP -s=4 [-bsp_sync_type=cdd]
               [-bsp_exch_strategy=fused] [-b]

Same flags (``bin/tmlauncher:31-64``) and the same session-cfg format: a bash-style
fragment assigning ``RULE MODELFILE MODELCLASS DEVICES [BSP_SYNC_TYPE BSP_EXCH_STRATEGY]``
where ``DEVICES`` is ``cuda0,cuda1`` or a bash array ``("node0:cuda0" "node1:cuda0,cuda1")``
(``:75-84,259-293``).  Device names are validated (``cuda[0-9]+``, ``:164-172``; ``cpuN`` is
additionally accepted for gloo runs), ``-b`` derives a CPU list per GPU from
``nvidia-smi topo -m`` (``:203-248``), BSP with one device runs ``avg``, multi-host BSP falls
back to a NCCL/host strategy (``:332-360``).
"""
from __future__ import annotations

import re
import shlex
import socket
import sys

RED, NC = "\033[1;31m", "\033[0m"


def err(msg):
    sys.stderr.write("%sError%s: %s\n" % (RED, NC, msg))
    raise SystemExit(1)


def parse_cfg(path):
    """Parse the bash-fragment session config without executing it."""
    text = open(path).read()
    text = re.sub(r"(?m)#.*$", "", text)
    out = {}
    # arrays: NAME=( ... ) possibly multi-line
    for m in re.finditer(r"(?s)\b([A-Z_]+)=\((.*?)\)", text):
        out[m.group(1)] = [s for s in shlex.split(m.group(2)) if s]
    text = re.sub(r"(?s)\b[A-Z_]+=\(.*?\)", "", text)
    for m in re.finditer(r"(?m)^\s*([A-Z_]+)=(.*)$", text):
        v = m.group(2).strip()
        out[m.group(1)] = shlex.split(v)[0] if v else ""
    return out


def check_device_name(d):
    if not re.match(r"^(cuda|cpu)[0-9]+$", d):
        err("device name should look like cuda0 (got %r)" % d)


def expand_devices(devices):
    """DEVICES → list of 'host:dev' strings (ref ``get_cpu_dev_array``, ``:259-293``)."""
    if isinstance(devices, str):
        devices = [devices]
    if not devices or not devices[0]:
        err("DEVICES empty")
    res = []
    if len(devices) == 1 and ":" not in devices[0]:
        host = socket.gethostname().split(".")[0]
        for d in devices[0].split(","):
            check_device_name(d)
            res.append("%s:%s" % (host, d))
        return res
    for s in devices:
        if ":" not in s:
            err("multi-host DEVICES entries must look like host:cuda0,cuda1 (got %r)" % s)
        host, ds = s.split(":", 1)
        for d in ds.split(","):
            check_device_name(d)
            res.append("%s:%s" % (host, d))
    return res


def parse_args(argv):
    opt = {}
    for a in argv:
        if a.startswith(("-cfg=", "--config=")):
            opt["CONFIG"] = a.split("=", 1)[1]
        elif a.startswith(("-file=", "--modelfile=")):
            opt["MODELFILE"] = a.split("=", 1)[1]
        elif a.startswith(("-class=", "--modelclass=")):
            opt["MODELCLASS"] = a.split("=", 1)[1]
        elif a.startswith(("-r=", "--rule=")):
            opt["RULE"] = a.split("=", 1)[1]
        elif a.startswith(("-s=", "--size=")):
            opt["SIZE"] = a.split("=", 1)[1]
        elif a.startswith("-bsp_sync_type="):
            opt["BSP_SYNC_TYPE"] = a.split("=", 1)[1]
        elif a.startswith("-bsp_exch_strategy="):
            opt["BSP_EXCH_STRATEGY"] = a.split("=", 1)[1]
        elif a.startswith(("-d=", "--devices=")):
            opt["DEVICES"] = a.split("=", 1)[1]
        elif a.startswith("-resume="):
            opt["RESUME"] = a.split("=", 1)[1]
        elif a in ("-b", "-bind"):
            opt["BIND"] = True
        elif a in ("-n", "--dry-run"):
            opt["DRY"] = True
        else:
            err("unknown option %s" % a)
    return opt


def resolve(opt):
    if not any(k in opt for k in ("CONFIG", "MODELFILE", "MODELCLASS", "RULE")):
        err("Neither config nor run options provided")
    if "CONFIG" in opt:
        cfg = parse_cfg(opt["CONFIG"])
        print("\nconfig file provided:")
        for k in ("RULE", "MODELFILE", "MODELCLASS", "DEVICES", "BSP_SYNC_TYPE", "BSP_EXCH_STRATEGY"):
            if k in cfg:
                opt[k] = cfg[k]            # the config overrides the command line (ref :80-84)
        for k in ("RULE", "MODELFILE", "MODELCLASS"):
            print("%s = %s" % (k, opt.get(k)))
    for k, msg in (("MODELFILE", "NO modelfile provided"), ("MODELCLASS", "NO modelclass provided"), ("RULE", "NO rule provided")):
        if not opt.get(k):
            err(msg)
    if "DEVICES" not in opt:
        n = int(opt.get("SIZE", 1))
        opt["DEVICES"] = ",".join("cuda%d" % i for i in range(n))
    devs = expand_devices(opt["DEVICES"])
    me = socket.gethostname().split(".")[0]
    plan = dict(rule=opt["RULE"].upper(), modelfile=opt["MODELFILE"], modelclass=opt["MODELCLASS"],
                devices=[d.split(":", 1)[1] if d.split(":", 1)[0] == me else d for d in devs],
                sync_type=opt.get("BSP_SYNC_TYPE"), exch_strategy=opt.get("BSP_EXCH_STRATEGY"),
                bind=bool(opt.get("BIND")), resume=opt.get("RESUME"))
    plan["cpulists"] = None
    if plan["bind"]:
        from .parallel.hwloc_utils import gpu_cpu_affinity, range_expand
        cache, cpul = {}, []
        for d in devs:
            host, dev = d.split(":", 1)
            if host not in cache:
                cache[host] = gpu_cpu_affinity(None if host == me else host)
            aff = cache[host].get(int(re.sub(r"\D", "", dev)))
            cpul.append(range_expand(aff) if aff else "")
        plan["cpulists"] = cpul
    return plan


def main(argv=None):
    opt = parse_args(list(sys.argv[1:] if argv is None else argv))
    plan = resolve(opt)
    import theanompi_b200 as tm
    rules = {"BSP": tm.BSP, "EASGD": tm.EASGD, "GOSGD": tm.GOSGD, "ASGD": tm.ASGD}
    if plan["rule"] not in rules:
        err("rule must be one of %s" % sorted(rules))
    if plan["rule"] == "BSP":
        if plan["sync_type"]:
            tm.BSP.sync_type = plan["sync_type"]
        if plan["exch_strategy"]:
            tm.BSP.exch_strategy = plan["exch_strategy"]
    rule = rules[plan["rule"]]()
    if plan["resume"]:
        rule.model_config["resume"] = plan["resume"]
    print("launching %s on %s" % (plan["rule"], plan["devices"]))
    if opt.get("DRY"):
        print(plan)
        return 0
    rule.init(plan["devices"], plan["modelfile"], plan["modelclass"], cpulists=plan["cpulists"])
    rc = rule.wait()
    return rc or 0


if __name__ == "__main__":
    sys.exit(main())
