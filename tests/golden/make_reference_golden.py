"""Copies the reference's own golden values for the bccFe regression case into tests/golden/reference_bccfe_ref.json
(run in the build container, where /root/reference is mounted; the GPU box never reads /root/reference).

Source: /root/reference/tests/scf/references/Example_bulk_bccFe_*/ref.json (totaldos.out rows, written by the reference's
Fortran program) and /root/reference/tests/scf/cases.json (the namelist patches that define each case)."""
import glob
import json
import os

REF = "/root/reference/tests/scf"
out = {"_source": "rslmtoasa/rslmtoasa tests/scf/references/Example_bulk_bccFe_*/ref.json + tests/scf/cases.json; tests/postproc/references/Example_exchange_{conductivity_fccPt,bccFe}*/ref.json + tests/postproc/cases.json", "cases": {}}
cases = {c["name"]: c for c in json.load(open(os.path.join(REF, "cases.json")))["cases"]}
for d in sorted(glob.glob(os.path.join(REF, "references", "Example_bulk_bccFe_*"))):
    name = os.path.basename(d)
    ref = json.load(open(os.path.join(d, "ref.json")))
    out["cases"][name] = {"namelists": cases[name]["namelists"], "totaldos.out": ref["text"]["totaldos.out"],
                          "etot": ref["nml"]["Fe_out.nml"]["etot"]}
# post-processing fixtures (conductivity): tests/postproc/references/Example_exchange_conductivity_fccPt*/ref.json
PP = "/root/reference/tests/postproc"
pcases = {c["name"]: c for c in json.load(open(os.path.join(PP, "cases.json")))["cases"]}
for d in sorted(glob.glob(os.path.join(PP, "references", "Example_exchange_conductivity_fccPt*"))):
    name = os.path.basename(d)
    out["cases"][name] = {"namelists": pcases[name]["namelists"], "Pt_cond.out": json.load(open(os.path.join(d, "ref.json")))["text"]["Pt_cond.out"]}
for d in sorted(glob.glob(os.path.join(PP, "references", "Example_exchange_bccFe*"))):
    name = os.path.basename(d)
    out["cases"][name] = {"namelists": pcases[name]["namelists"], "jij.out": json.load(open(os.path.join(d, "ref.json")))["text"]["jij.out"]}
here = os.path.dirname(os.path.abspath(__file__))
json.dump(out, open(os.path.join(here, "reference_bccfe_ref.json"), "w"), indent=1, sort_keys=True)
print(len(out["cases"]), "cases")
