"""Darknet cfg reader with the semantics of the reference's utils/parse_config.py:1-18."""


def parse_model_config(path):
    """Return the list of blocks of a yolo-v3 cfg: [{'type': ..., key: value, ...}, ...].

    Lines that are empty or start with '#' are skipped (tested before whitespace is stripped, as the
    reference does); every [convolutional] block starts with batch_normalize=0; keys and values are
    split on '=' and stripped."""
    with open(path, "r") as f:
        raw = f.read().split("\n")
    module_defs = []
    for text in raw:
        if not text or text.startswith("#"):
            continue
        text = text.strip()
        if text.startswith("["):
            block = {"type": text[1:-1].rstrip()}
            if block["type"] == "convolutional":
                block["batch_normalize"] = 0
            module_defs.append(block)
        else:
            key, value = text.split("=")
            module_defs[-1][key.rstrip()] = value.strip()
    return module_defs
