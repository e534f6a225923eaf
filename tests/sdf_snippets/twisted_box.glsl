// twist: a rotation whose angle depends on the height, built as a mat2 variable
float sdf(in vec3 p) {
    const float k = 2.5;
    float c = cos(k * p.y);
    float s = sin(k * p.y);
    mat2 m = mat2(c, -s, s, c);
    vec3 q = vec3(m * p.xz, p.y);
    vec3 d = abs(q) - vec3(0.35, 0.2, 0.6);
    return (length(max(d, 0.0)) + min(max(d.x, max(d.y, d.z)), 0.0)) * 0.6;
}

float sdfmaterial(in vec3 p) {
    return 2.0;
}
