// the classic polar-coordinate mandelbulb distance estimator (atan / acos / pow)
float sdf(in vec3 p) {
    vec3 z = p;
    float dr = 1.0;
    float r = 0.0;
    const float power = 8.0;
    for (int i = 0; i < 6; i++) {
        r = length(z);
        if (r > 2.0) break;
        float theta = acos(z.z / r);
        float phi = atan(z.y, z.x);
        dr = pow(r, power - 1.0) * power * dr + 1.0;
        float zr = pow(r, power);
        theta = theta * power;
        phi = phi * power;
        z = zr * vec3(sin(theta) * cos(phi), sin(phi) * sin(theta), cos(theta));
        z += p;
    }
    return 0.5 * log(r) * r / dr;
}

float sdfmaterial(in vec3 p) {
    return mix(1.0, 2.0, clamp(length(p), 0.0, 1.0));
}
