"""metadata.json contract of PRISMA (reference: bands/common/meta.py:14-146, README.md:70-115), re-implemented.

Same function names and semantics as the reference module so band scripts read the same way:
a PRISMA folder holds `metadata.json` = {"bands": {<band>: {"url": file, ["folder": sub], ["values": {...}]}}, ...}.
"""
import json
import os

META_FILE = "metadata.json"


def get_metadata_path(path):
    if os.path.isfile(path):
        return path if path.endswith(".json") else get_metadata_path(os.path.dirname(path))
    if os.path.isdir(path):
        return os.path.join(path, META_FILE)
    return None


def load_metadata(path):
    mp = get_metadata_path(path)
    if mp and os.path.exists(mp):
        with open(mp) as f:
            return json.load(f)
    return None


def write_metadata(path, metadata):
    if metadata is None:
        return
    mp = get_metadata_path(path)
    if mp and os.path.exists(mp):
        with open(mp, "w") as f:
            f.write(json.dumps(metadata, indent=4))


def create_metadata(path):
    folder = os.path.dirname(path) if os.path.isfile(path) else path
    os.makedirs(folder, exist_ok=True)
    mp = os.path.join(folder, META_FILE)
    if not os.path.exists(mp):
        with open(mp, "w") as f:
            f.write(json.dumps({"bands": {}}, indent=4))
    return load_metadata(mp)


def is_video(path):
    return path.endswith(".mp4")


def add_band(metadata, band, url="", folder=""):
    entry = metadata.setdefault("bands", {}).setdefault(band, {})
    if url != "":
        entry["url"] = url
    if folder != "":
        entry["folder"] = folder


def get_url(path, metadata, band):
    if os.path.isdir(path) and metadata:
        url = metadata.get("bands", {}).get(band, {}).get("url")
        if url:
            return os.path.join(path, url)
    return path


def get_target(path, metadata, band="rgba", target="", force_extension=None):
    folder = target if os.path.isdir(target) else os.path.dirname(path)
    ext = os.path.basename(path).rsplit(".", 1)[1]
    if force_extension and (not is_video(path) or force_extension == "csv"):
        ext = force_extension
    name = band + "." + ext
    if target == "" or os.path.isdir(target):
        target = os.path.join(folder, name)
    if metadata:
        add_band(metadata, band, url=name)
    return target


def set_default_band(path, band, band_default):
    data = load_metadata(path)
    if data and band_default in data.get("bands", {}):
        data["bands"][band] = data["bands"][band_default]
        write_metadata(path, data)
