def named_apply(fn, module, name="", depth_first=True, include_root=False):
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = ".".join((name, child_name)) if name else child_name
        named_apply(fn, child, child_name, depth_first, True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module
