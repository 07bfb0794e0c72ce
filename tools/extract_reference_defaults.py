"""One-off generator of cupix_b200/data/colore_arinyo_fits.json.

The reference reads its CoLoRe best-fit Arinyo values from FITS headers with astropy
(cupix/likelihood/model_lya.py:96-111).  astropy is not a dependency of this package, so the
eight header values per redshift bin are extracted once, here, with a plain 80-byte-card parser
and committed as JSON.  Run inside the build container only (needs /root/reference).
"""
import json
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
KEYS = ["bias_LYA", "beta_LYA", "dnl_arinyo_q1", "dnl_arinyo_q2", "dnl_arinyo_kv",
        "dnl_arinyo_av", "dnl_arinyo_bv", "dnl_arinyo_kp"]

out = {}
for z in ("2.2", "2.4", "2.6", "2.8"):
    raw = open(f"{REF}/data/colore_xi/bin_{z}/lyaxlya.fits", "rb").read()
    # header of HDU 1 follows the primary HDU; just scan the first blocks for the cards we need
    cards = [raw[i:i + 80].decode("ascii", "replace") for i in range(0, min(len(raw), 2880 * 64), 80)]
    vals = {}
    for c in cards:
        m = re.match(r"^(HIERARCH\s+)?([A-Za-z0-9_ ]+?)\s*=\s*([^/]+)", c)
        if m and m.group(2).strip() in KEYS and m.group(2).strip() not in vals:
            vals[m.group(2).strip()] = float(m.group(3).strip())
    out[z] = vals
json.dump(out, open("cupix_b200/data/colore_arinyo_fits.json", "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
