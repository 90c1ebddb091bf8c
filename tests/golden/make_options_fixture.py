"""Extracts the default of every leaf option of the reference's share/xtp/xml/subpackages/gwbse.xml and the
gwbse subtrees of its integration-test option files into tests/golden/gwbse_xml_options.json (run in the build
container, where /root/reference exists)."""
import glob
import json
import os
import xml.etree.ElementTree as ET

REF = "/root/reference/xtp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gwbse_xml_options.json")


def leaves(node, prefix=""):
    for child in node:
        key = prefix + child.tag
        if len(child):
            yield from leaves(child, key + ".")
        else:
            yield key, child


def main():
    root = ET.parse(os.path.join(REF, "share/xtp/xml/subpackages/gwbse.xml")).getroot()
    defaults = {k: n.get("default") for k, n in leaves(root) if n.get("default") is not None}
    files = {}
    for path in sorted(glob.glob(os.path.join(REF, "src/tests/DataFiles/xtp_tools_integration_tests/dftgwbse_*.xml"))):
        gw = ET.parse(path).getroot().find("dftgwbse/gwbse")
        files[os.path.basename(path)] = {k: (n.text or "").strip() for k, n in leaves(gw)}
    with open(OUT, "w") as fh:
        json.dump({"defaults": defaults, "files": files}, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
