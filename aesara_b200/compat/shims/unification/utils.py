def transitive_get(key, d):
    """Follow ``key -> d[key] -> d[d[key]] ...`` until a value that is not a
    key of ``d`` (the "walk" of a substitution)."""
    while True:
        try:
            if key in d:
                key = d[key]
            else:
                return key
        except TypeError:  # unhashable
            return key
