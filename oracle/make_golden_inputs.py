"""Random multifurcating superdtl inputs for the tests (TEST INFRASTRUCTURE; no reference needed):
random-join trees where a join takes 3-5 roots with some probability, so polytomies nest and can sit at the
root; only leaves are named, as in data/crispr-class1.in.json."""

COST_NAMES = ["SPECIATION", "DUPLICATION", "HORIZONTAL_TRANSFER", "FULL_LOSS", "SEGMENTAL_LOSS"]


def _random_tree(rng, leaves, polytomy):
    roots = list(leaves)
    while len(roots) > 1:
        arity = 2
        if len(roots) >= 3 and rng.random() < polytomy:
            arity = min(len(roots), rng.choice([3, 3, 4, 5]))
        picked = [roots.pop(rng.randrange(len(roots))) for _ in range(arity)]
        roots.append("(" + ",".join(picked) + ")")
    return roots[0] + ";"


def random_polytomy_instance(rng, n_species, n_objects, families):
    species_leaves = [f"S{i}" for i in range(n_species)]
    object_leaves = [f"g{i}" for i in range(n_objects)]
    names = [f"f{i}" for i in range(families)]
    syntenies = {}
    for leaf in object_leaves:
        keep = [f for f in names if rng.random() < 0.6] or [rng.choice(names)]
        syntenies[leaf] = keep
    return {
        "object_tree": _random_tree(rng, object_leaves, 0.45),
        "species_tree": _random_tree(rng, species_leaves, 0.0),
        "leaf_object_species": {leaf: rng.choice(species_leaves) for leaf in object_leaves},
        "leaf_syntenies": syntenies,
        "costs": dict(zip(COST_NAMES, rng.choice([[0, 1, 1, 1, 1], [1, 2, 1, 1, 2], [0, 1, 2, 1, 3], [2, 0, 1, 0, 1]]))),
    }
